"""Developer tool (one GPU): per-stage times of ONE tile-row shard of a frame (rank r of N, work-balanced rows, shard
cull on) next to the whole frame -- what a rank of the tile-row split spends where.  usage: shard_stage.py [workload] [N]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
import sharding  # noqa: E402
from diff_gaussian_rasterization import _C  # noqa: E402
from renderer import FrameRenderer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload(name)
W, H = w["W"], w["H"]
gy = (H + 15) // 16
fr = FrameRenderer(cloud, W, H, [1.0, 1.0, 1.0], dev)
v = fr.upload_view(views[1])
fr.render(v)
scene = fr._scene(v, None)
ncon = _C.fetch("n_contrib", scene, fr.geom, fr.binning, fr.img, fr.capacity).view(H, W).to(dev)
_b, need = bench.algorithmic_blend_bytes(ncon, W, H)
rng = _C.fetch("ranges", scene, fr.geom, fr.binning, fr.img, fr.capacity).view(-1, 2).to(torch.int64)
inst = (rng[:, 1] - rng[:, 0]).view(gy, -1)
rows = sharding.balanced_rows(sharding.row_cost(need.cpu().numpy(), inst.numpy()), N)
L = _C.lib()


def stages(tile_rows, cull):
    L.gs_profile_enable(1)
    ms4, tot = np.zeros(4, dtype=np.float32), np.zeros(4)
    n = 12
    for i in range(n + 2):
        fr.enqueue(v, tile_rows=tile_rows, shard_cull=cull)
        L.gs_profile_read(ms4.ctypes.data)
        if i >= 2:
            tot += ms4
    L.gs_profile_enable(0)
    ms = bench._median_ms(lambda i: fr.enqueue(v, tile_rows=tile_rows, shard_cull=cull), n, warm=2)
    return {"frame_ms": round(ms, 4), "cull+preprocess/sort/lists/blend": [round(float(x), 4) for x in tot / n]}


print(json.dumps({"whole frame": stages(None, False)}))
for r in sorted({0, N // 2, N - 1}):
    print(json.dumps({f"rank {r} of {N}, rows {rows[r]}": stages(tuple(rows[r]), True)}))
