"""Developer tool (GPU box): SURVEY 8f-4 at C2 size.  (a) gs_decode_head against the reference's torch expressions
(models/model_v2.py:287-375 restated in oracle/head.py, run on the GPU); (b) a C2 frame from the padded (P,13,3) SH
array with sh_degree 1 against the packed (P,1,3) array with sh_degree 0 (same image), one stream and six frames in
flight.  Writes gpurun_out/bench_head.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
from diff_gaussian_rasterization import _C  # noqa: E402
from oracle import head  # noqa: E402
from renderer import FramePipeline, FrameRenderer  # noqa: E402

dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload("C2")
W, H, P = w["W"], w["H"], cloud["means3D"].shape[0]
rng = np.random.default_rng(0)
feat = torch.from_numpy((0.5 * rng.standard_normal((P, 8))).astype(np.float32)).to(dev)
rgb = torch.rand(P, 3, device=dev)
prim = torch.randint(0, 1024, (P, 3), device=dev).float()


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ours = lambda: _C.decode_head(feat, rgb, prim, scale_factor=448, xyz_offset=512)
def theirs():  # the restated torch expressions on CUDA tensors (the oracle's constants are created on the default device)
    with torch.device(dev):
        return head.decode_head(feat, rgb, prim, scale_factor=448, xyz_offset=512)  # torch ops on CUDA tensors
a, b = ours(), theirs()
same = all(torch.equal(a[k], b[k][:, :1] if k == "shs" else b[k]) for k in ("means3D", "rotations", "scales", "opacities", "shs"))
out = {"workload": f"head decode, P={P}, 8 feature columns", "bit_identical_to_torch_cuda": bool(same),
       "gs_decode_head_ms": timeit(ours), "torch_expressions_ms": timeit(theirs)}

packed = dict(cloud, shs=cloud["shs"][:, :1].contiguous(), sh_degree=0)
res = {}
for name, cl in (("padded_M13_deg1", cloud), ("packed_M1_deg0", packed)):
    fr = FrameRenderer(cl, W, H, [1.0, 1.0, 1.0], dev, capacity=24_000_000)
    vd = [fr.upload_view(v) for v in views]
    img = fr.enqueue(vd[3]).clone()
    one = timeit(lambda: fr.enqueue(vd[7]), n=100)
    pipe = FramePipeline(cl, W, H, [1.0, 1.0, 1.0], dev, depth=6, capacity=24_000_000)

    def go(m, off):
        pipe.begin()
        for i in range(m):
            pipe.enqueue(vd[(off + i) % len(vd)], slot=i)
        pipe.end()

    go(30, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    go(600, 30)
    e1.record()
    torch.cuda.synchronize()
    res[name] = {"one_frame_ms": one, "frames_per_s_6_in_flight": 600 / (e0.elapsed_time(e1) / 1e3), "img": img}
out["same_image"] = bool(torch.equal(res["padded_M13_deg1"].pop("img"), res["packed_M1_deg0"].pop("img")))
out["frames"] = res
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_head.json"), "w"), indent=1)
print(json.dumps(out))
