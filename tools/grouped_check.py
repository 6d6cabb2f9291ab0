"""Developer tool (GPU box): the grouped hit loop of the blend (GSPLAT_B200_BLEND_GROUPED=1) against the plain loop --
bit-exactness of image / final_T / n_contrib, single-frame and blend-stage time, and throughput with frames in flight.
usage: grouped_check.py [workload]"""
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
from renderer import FrameRenderer, FramePipeline  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload(name)
W, H = w["W"], w["H"]
cap = 40_000_000 if name == "C4" else 24_000_000
fr = FrameRenderer(cloud, W, H, [1.0, 1.0, 1.0], dev, capacity=cap)
vd = [fr.upload_view(v) for v in views]
L = _C.lib()
probe = sorted(set([0, len(vd) // 5, len(vd) // 2, (3 * len(vd)) // 4]))


def snap(k):
    img = fr.render(vd[k]).clone()
    sc = fr._scene(vd[k], None)
    return (img, _C.fetch("final_T", sc, fr.geom, fr.binning, fr.img, fr.capacity).clone(),
            _C.fetch("n_contrib", sc, fr.geom, fr.binning, fr.img, fr.capacity).clone())


def setmode(grouped, full, hi=0, longsms=0, ctas=0):
    raise RuntimeError("the library reads GSPLAT_B200_BLEND_PLAIN once per process; run this tool once per mode")


setmode(0, 0)
exact = {k: snap(k) for k in probe}
pipe = FramePipeline(cloud, W, H, [1.0, 1.0, 1.0], dev, depth=6 if name != "C4" else 4, capacity=cap)
pvd = [pipe.upload_view(v) for v in views]
MODES = [tuple(int(x) for x in m.split(",")) for m in sys.argv[2:]] or [(0, 0, 0), (1, 0, 0), (1, 0, 1), (1, 0, 2), (1, 1, 1)]
for mode in MODES:
    grouped, full, hi = mode[:3]
    longsms = mode[3] if len(mode) > 3 else 0
    ctas = mode[4] if len(mode) > 4 else 0
    setmode(grouped, full, hi, longsms, ctas)
    same = True
    for k in probe:
        got = snap(k)
        same = same and all(bool(torch.equal(a, b)) for a, b in zip(got, exact[k]))
    L.gs_profile_enable(1)
    ms4, tot = np.zeros(4, dtype=np.float32), np.zeros(4)
    n = min(len(vd), 40)
    for i in range(n):
        fr.enqueue(vd[(3 * i) % len(vd)])
        L.gs_profile_read(ms4.ctypes.data)
        tot += ms4
    L.gs_profile_enable(0)
    frame_ms = bench._median_ms(lambda i: fr.enqueue(vd[(3 * i) % len(vd)]), n, warm=2)
    # throughput with frames in flight
    def run(nf):
        pipe.begin()
        for i in range(nf):
            pipe.enqueue(pvd[(i * 7) % len(pvd)])
        pipe.end()
    run(24)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nf = 480 if name != "C4" else 96
    e0.record()
    run(nf)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"grouped": grouped, "fullgrid": full, "himode": hi, "long_sms": longsms, "ctas": ctas, "bit_identical": same, "frame_ms": round(frame_ms, 4),
                      "stage_ms": [round(float(x), 4) for x in tot / n],
                      "in_flight_fps": round(nf / (e0.elapsed_time(e1) / 1e3), 1)}), flush=True)
