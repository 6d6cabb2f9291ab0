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
        fr.enqueue(vd, tile_rows=(r0, r1), peer_out=ptrs)
    barrier()
    torch.cuda.synchronize()
    same = bool(torch.equal(img, full))
    ok = ok and same
    if rank == 0:
        print(f"view {i}: rows {rows} assembled == full: {same}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER CHECK", "OK" if int(flag) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
