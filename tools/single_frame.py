"""Developer tool (GPU box): one C2 frame at a time (latency mode) -- per-stage times from the library's events and the
median frame time over the orbit.  Prints one JSON line."""
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
from renderer import FrameRenderer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload(name)
fr = FrameRenderer(cloud, w["W"], w["H"], [1.0, 1.0, 1.0], dev, capacity=40_000_000 if name == "C4" else 24_000_000)
vd = [fr.upload_view(v) for v in views]
L = _C.lib()
for v in vd[:3]:
    fr.render(v)
L.gs_profile_enable(1)
ms4, tot = np.zeros(4, dtype=np.float32), np.zeros(4)
n = min(len(vd), 40)
for i in range(n):
    fr.enqueue(vd[(3 * i) % len(vd)])
    L.gs_profile_read(ms4.ctypes.data)
    tot += ms4
L.gs_profile_enable(0)
frame_ms = bench._median_ms(lambda i: fr.enqueue(vd[(3 * i) % len(vd)]), n, warm=2)
print(json.dumps({"frame_ms": round(frame_ms, 4), "stage_ms": [round(float(x), 4) for x in tot / n]}))
