"""Developer tool (GPU box): the blend's split-walk latency mode (GsScene.blend_split) against the exact frame --
largest pixel / final_T difference, share of pixels whose contributor count differs, and the single-frame / blend-stage
time for several hand-over thresholds.  usage: split_check.py [workload] [threshold ...]"""
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
thresholds = [int(x) for x in sys.argv[2:]] or [0, 8, 16, 32, 64]
dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload(name)
W, H = w["W"], w["H"]
cap = 40_000_000 if name == "C4" else 24_000_000
ref = FrameRenderer(cloud, W, H, [1.0, 1.0, 1.0], dev, capacity=cap)
vd = [ref.upload_view(v) for v in views]
L = _C.lib()
probe = [0, len(vd) // 5, len(vd) // 2, (3 * len(vd)) // 4]
exact = {}
for k in probe:
    img = ref.render(vd[k]).clone()
    sc = ref._scene(vd[k], None)
    exact[k] = (img, _C.fetch("final_T", sc, ref.geom, ref.binning, ref.img, ref.capacity).clone(),
                _C.fetch("n_contrib", sc, ref.geom, ref.binning, ref.img, ref.capacity).clone())
for t in thresholds:
    fr = FrameRenderer(cloud, W, H, [1.0, 1.0, 1.0], dev, capacity=cap, share=ref, blend_split=t)
    dmax, tmax, nfrac = 0.0, 0.0, 0.0
    for k in probe:
        img = fr.render(vd[k])
        sc = fr._scene(vd[k], None)
        T = _C.fetch("final_T", sc, fr.geom, fr.binning, fr.img, fr.capacity)
        n = _C.fetch("n_contrib", sc, fr.geom, fr.binning, fr.img, fr.capacity)
        dmax = max(dmax, float((img - exact[k][0]).abs().max()))
        tmax = max(tmax, float((T - exact[k][1]).abs().max()))
        nfrac = max(nfrac, float((n != exact[k][2]).float().mean()))
    L.gs_profile_enable(1)
    ms4, tot = np.zeros(4, dtype=np.float32), np.zeros(4)
    n = min(len(vd), 40)
    for i in range(n):
        fr.enqueue(vd[(3 * i) % len(vd)])
        L.gs_profile_read(ms4.ctypes.data)
        tot += ms4
    L.gs_profile_enable(0)
    frame_ms = bench._median_ms(lambda i: fr.enqueue(vd[(3 * i) % len(vd)]), n, warm=2)
    print(json.dumps({"blend_split": t, "max_abs_pixel_diff": dmax, "max_abs_T_diff": tmax,
                      "n_contrib_mismatch_share": nfrac, "frame_ms": round(frame_ms, 4),
                      "stage_ms": [round(float(x), 4) for x in tot / n]}), flush=True)
