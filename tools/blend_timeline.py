"""Developer tool (GPU box): per-unit timeline of the blend-forward kernel (needs `make -C .../csrc timeline`).
Writes gpurun_out/blend_timeline.npz and prints a summary.  usage: GSPLAT_B200_LIB=.../libgsplat_b200_tl.so python tools/blend_timeline.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
from diff_gaussian_rasterization import _C  # noqa: E402
from renderer import FrameRenderer  # noqa: E402

dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload("C2")
fr = FrameRenderer(cloud, w["W"], w["H"], [1.0, 1.0, 1.0], dev, capacity=24_000_000)
L = _C.lib()
units = ((w["W"] + 15) // 16) * ((w["H"] + 15) // 16) * 8
buf = torch.zeros(units * 12, dtype=torch.int64, device=dev)
vd = fr.upload_view(views[7])
for _ in range(3):
    fr.render(vd)
L.gs_debug_timeline.argtypes = [C.c_void_p]
assert L.gs_debug_timeline(buf.data_ptr()) == 0
fr.render(vd)
t = buf.cpu().numpy().reshape(-1, 12)
t = t[t[:, 0] != 0]
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3          # the warp-per-block part (until done or parked)
parked = t[:, 6] != 0
tstart, tend = (t[:, 6] - t0) / 1e3, (t[:, 7] - t0) / 1e3        # the team part of parked blocks
tbatches, thits = t[:, 8] >> 32, t[:, 8] & 0xffffffff
fin = np.where(parked, tend, end)
sm = t[:, 2] >> 32
total = t[:, 2] & 0xffffffff
batches, hits = t[:, 3] >> 32, t[:, 3] & 0xffffffff
dur = end - start
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "blend_timeline.npz"), t=t, start=start, end=end)
print(f"units with work {len(t)}  kernel span {fin.max():.1f} us  sum(dur) {dur.sum()/1e3:.2f} ms  parked {parked.sum()}"
      f"  team time sum {(tend - tstart)[parked].sum()/1e3:.2f} ms")
print("longest units (us):")
for i in np.argsort(-fin)[:12]:
    print(f"  start {start[i]:7.1f} warp-end {end[i]:7.1f} batches {batches[i]:5d} hits {hits[i]:6d} list {total[i]:6d} sm {sm[i]}"
          + (f" | team {tstart[i]:7.1f} -> {tend[i]:7.1f} ({tend[i]-tstart[i]:6.1f} us) batches {tbatches[i]:5d} hits {thits[i]:6d}"
             + f" | consumer wait {t[i,9]/1965:6.1f} us | producer0: total {t[i,5]/1965:6.1f} rec {(t[i,10]>>32)/1965:5.1f} tok {(t[i,10]&0xffffffff)/1965:5.1f} room {(t[i,11]>>32)/1965:5.1f} eval {(t[i,11]&0xffffffff)/1965:5.1f}"
             if parked[i] else ""))
if parked.any():
    td = (tend - tstart)[parked]
    wait = (tstart - end)[parked]
    print(f"team blocks: n {parked.sum()} dur mean {td.mean():.1f} median {np.median(td):.1f} max {td.max():.1f} us; wait for a team "
          f"mean {wait.mean():.1f} max {wait.max():.1f} us; us per team hit {td.sum()/max(1,thits[parked].sum()):.4f}; "
          f"us per team batch {td.sum()/max(1,tbatches[parked].sum()):.3f}")
    A = np.stack([tbatches[parked], thits[parked], np.ones(parked.sum())], 1).astype(np.float64)
    coef, *_ = np.linalg.lstsq(A, td, rcond=None)
    print("team dur ~ %.4f us*batches + %.4f us*hits + %.3f us" % tuple(coef))
for q in (50, 75, 90, 95, 99, 100):
    print(f"  time by which {q}% of the units are finished: {np.percentile(fin, q):.1f} us")
edges = np.linspace(0, fin.max(), 21)
act = [(np.minimum(end, b) - np.maximum(start, a)).clip(0).sum() / (b - a) for a, b in zip(edges[:-1], edges[1:])]
tact = [(np.minimum(tend, b) - np.maximum(tstart, a)).clip(0)[parked].sum() / (b - a) for a, b in zip(edges[:-1], edges[1:])]
print("active block warps per 5% time slice:", " ".join(f"{a:.0f}" for a in act))
print("active teams per 5% time slice:", " ".join(f"{a:.0f}" for a in tact))
A = np.stack([batches, hits, np.ones_like(hits)], 1).astype(np.float64)
coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
print("warp dur ~ %.4f us*batches + %.4f us*hits + %.3f us" % tuple(coef))
