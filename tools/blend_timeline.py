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
start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
sm = t[:, 2] >> 32
total = t[:, 2] & 0xffffffff
batches, hits = t[:, 3] >> 32, t[:, 3] & 0xffffffff
dur = end - start
wait_us, loop_us = (t[:, 4] & 0xffffff) / 1965.0, t[:, 5] / 1965.0  # SM clock 1965 MHz
entry, entry_live, final_live = t[:, 4] >> 40, (t[:, 4] >> 32) & 0xff, (t[:, 4] >> 24) & 0xff
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "blend_timeline.npz"), start=start, end=end, sm=sm, total=total,
                    batches=batches, hits=hits, entry=entry, entry_live=entry_live, final_live=final_live, loop_us=loop_us)
print(f"units with work {len(t)}  kernel span {end.max():.1f} us  sum(dur) {dur.sum()/1e3:.2f} ms  -> mean concurrency {dur.sum()/end.max():.0f} warps")
print("longest units (us, start, batches, hits, list):")
for i in np.argsort(-dur)[:12]:
    print(f"  dur {dur[i]:7.1f} start {start[i]:7.1f} end {end[i]:7.1f} batches {batches[i]:5d} hits {hits[i]:6d} list {total[i]:6d} sm {sm[i]} wait {wait_us[i]:6.1f} hitloop {loop_us[i]:6.1f} tail@{entry[i]} live {entry_live[i]} final_live {final_live[i]} | cull+wait {t[i,8]/1965:.0f} eval {t[i,6]/1965:.0f} chain {t[i,7]/1965:.0f} row-eval {t[i,9]/1965:.0f} row-chain {t[i,10]/1965:.0f}")
long_units = batches > 64
print("units > 64 batches: %d, of which entered tail mode: %d; final live histogram (never-tail long units):" % (long_units.sum(), (long_units & (entry > 0)).sum()),
      np.histogram(final_live[long_units & (entry == 0)], bins=[0, 1, 9, 17, 25, 33, 41, 49, 57, 65])[0])
for q in (50, 75, 90, 95, 99, 100):
    print(f"  time by which {q}% of the unit-time is done: {np.percentile(end, q):.1f} us")
edges = np.linspace(0, end.max(), 21)
act = [(np.minimum(end, b) - np.maximum(start, a)).clip(0).sum() / (b - a) for a, b in zip(edges[:-1], edges[1:])]
print("active warps per 5% time slice:", " ".join(f"{a:.0f}" for a in act))
print("us per batch: %.3f  us per hit-iteration: model fit" % (dur.sum() / batches.sum()))
A = np.stack([batches, hits, np.ones_like(hits)], 1).astype(np.float64)
coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
print("dur ~ %.4f us*batches + %.4f us*hits + %.3f us" % tuple(coef))
