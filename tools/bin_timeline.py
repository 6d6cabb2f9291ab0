"""Developer tool (GPU box): per-CTA phase timeline of the binning kernels (needs `make -C .../csrc timeline`)."""
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
SLOTS = 4096 + 2 * 16384
buf = torch.zeros(SLOTS * 8, dtype=torch.int64, device=dev)
vd = fr.upload_view(views[7])
for _ in range(3):
    fr.render(vd)
L.gs_debug_bin_timeline.argtypes = [C.c_void_p]
assert L.gs_debug_bin_timeline(buf.data_ptr()) == 0
fr.render(vd)
t = buf.cpu().numpy().reshape(SLOTS, 8)


def report(name, rows, nph):
    rows = rows[rows[:, 0] != 0]
    if not len(rows):
        return
    t0 = rows[:, 0].min()
    r = (rows[:, :nph] - t0) / 1e3
    print(f"{name}: {len(rows)} chunks, span {r[:, nph-1].max():.1f} us; start skew p50/p99 {np.percentile(r[:,0],50):.1f}/{np.percentile(r[:,0],99):.1f}")
    d = np.diff(r, axis=1)
    print("   phase durations us (mean / max): " + "  ".join(f"{d[:,k].mean():.1f}/{d[:,k].max():.1f}" for k in range(nph - 1)))
    worst = np.argmax(r[:, nph - 1])
    print("   last-finishing chunk", worst, "marks", " ".join(f"{x:.1f}" for x in r[worst]))


for p in range(4):
    report(f"depth pass {p} [load+rank | scan | chain | reorder | write]", t[p * 1024:(p + 1) * 1024], 6)
report("row pass [load+masks | scan | chain | scatter]", t[4096:4096 + 16384], 5)
report("col pass [load+masks | scan | chain | scatter]", t[4096 + 16384:], 5)
