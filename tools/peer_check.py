"""Developer tool (GPU box, torchrun with N >= 2): tile-row sharded frames assembled by peer stores must equal the
frame rendered by one GPU, bit for bit.  usage: torchrun --nproc-per-node N tools/peer_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import scenes  # noqa: E402
import sharding  # noqa: E402
from renderer import FrameRenderer  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
W, H = 1000, 600
cl = scenes.human_cloud(120000, scale_factor=320.0, seed=3, opacity="uniform")
views = [scenes.make_view(c, W, H) for c in scenes.orbit_c2w(6)]
fr = FrameRenderer(cl, W, H, [1, 1, 1], dev, capacity=20_000_000)
peer = sharding.PeerFrame(H, W, dev)
gy = (H + 15) // 16
ok = True
for i, v in enumerate(views):
    vd = fr.upload_view(v)
    full = fr.render(vd).clone()                       # every rank renders the whole frame: the expected image
    cost = np.ones(gy) + np.arange(gy) % 3             # arbitrary uneven partition, identical on every rank
    rows = sharding.balanced_rows(cost, world)
    img, ptrs, barrier = peer.next()
    r0, r1 = rows[rank]
    if r1 > r0:
        fr.enqueue(vd, tile_rows=(r0, r1), peer_out=ptrs, shard_cull=bool(i % 2))  # with and without the shard cull
    barrier()
    torch.cuda.synchronize()
    same = bool(torch.equal(img, full))
    ok = ok and same
    if rank == 0:
        print(f"view {i}: rows {rows} assembled == full: {same}", flush=True)

# ---- sharded backward: blend stage per rank over its rows, ONE all-reduce of the partial arrays, per-Gaussian stage
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402

wgt = torch.from_numpy(np.random.default_rng(7).standard_normal((3, H, W)).astype(np.float32)).to(dev)
v = views[2]
vd = fr.upload_view(v)
rs = GaussianRasterizationSettings(H, W, v.tanfovx, v.tanfovy, torch.ones(3, device=dev), 1.0, vd[0], vd[1], 1, vd[2],
                                   False, False)
rows = sharding.balanced_rows(np.ones(gy) + np.arange(gy) % 3, world)


def grads(tile_rows, group):
    d = {k: cl[k].to(dev).requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
    color, _ = GaussianRasterizer(rs, tile_rows=tile_rows, grad_group=group)(
        d["means3D"], torch.zeros_like(d["means3D"], requires_grad=True), d["opacities"], shs=d["shs"],
        scales=d["scales"], rotations=d["rotations"])
    color.backward(wgt)
    return {k: x.grad for k, x in d.items()}


g_full = grads(None, None)
g_shard = grads(rows[rank], dist.group.WORLD)
torch.cuda.synchronize()
worst = max(float((g_shard[k] - g_full[k]).abs().max() / (g_full[k].abs().max() + 1e-30)) for k in g_full)
ok = ok and worst <= 1e-3
if rank == 0:
    print(f"sharded backward + all-reduce vs single-GPU backward: worst relative error {worst:.2e}", flush=True)

# ---- e2e fan-out: every rank uploads 1/N of the host cloud, NVLink all-gathers complete it; frames must not change
from renderer import FramePipeline  # noqa: E402

packed = scenes.pack_cloud(cl)
host = {k: packed[k].contiguous().pin_memory() for k in ("means3D", "opacities", "scales", "rotations", "shs")}
pipe = FramePipeline(packed, W, H, [1, 1, 1], dev, depth=3, capacity=20_000_000)
outs = [torch.empty((3, H, W), dtype=torch.float32).pin_memory() for _ in range(7)]
hv = [tuple(torch.from_numpy(a).pin_memory() for a in (v.viewmatrix, v.projmatrix, v.campos)) for v in views]
pipe.begin()
for i in range(7):
    k = (rank + i) % len(views)
    pipe.enqueue_host(host, hv[k], (views[k].tanfovx, views[k].tanfovy), outs[i], slot=i, group=dist.group.WORLD)
pipe.end()
torch.cuda.synchronize()
for i in range(7):
    k = (rank + i) % len(views)
    same = bool(torch.equal(outs[i].to(dev), fr.render(fr.upload_view(views[k]))))
    ok = ok and same
if rank == 0:
    print(f"fan-out frames == resident frames: {ok}", flush=True)

flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER CHECK", "OK" if int(flag) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
