"""Developer tool (GPU box): H2D bandwidth of a strided (2-D) copy of the SH coefficients a degree actually reads
(48 of 156 bytes per point at C2) against the full 1-D copy, pinned host memory, copy engine."""
import ctypes
import glob
import os
import sys

import torch

P, M, D = 799957, 13, 1
used = (D + 1) ** 2 * 12
pitch = M * 12
host = torch.empty(P * pitch, dtype=torch.uint8).pin_memory()
dev_full = torch.empty(P * pitch, dtype=torch.uint8, device="cuda")
dev_pack = torch.empty(P * used, dtype=torch.uint8, device="cuda")
paths = [p for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))]
rt = ctypes.CDLL(paths[0] if paths else "libcudart.so")
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                 ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_full = timeit(lambda: dev_full.copy_(host, non_blocking=True))
print(f"1-D full  {P*pitch/1e6:.1f} MB  {t_full:.3f} ms  {P*pitch/t_full/1e6:.1f} GB/s")
for dpitch in (used, pitch):
    dst = dev_pack if dpitch == used else dev_full
    t = timeit(lambda: rt.cudaMemcpy2DAsync(dst.data_ptr(), dpitch, host.data_ptr(), pitch, used, P, 1, st))
    print(f"2-D {used}B of {pitch}B rows -> dpitch {dpitch}: {t:.3f} ms  payload {P*used/t/1e6:.1f} GB/s")

# zero-copy gather kernel (tools/zc_test.cu)
zc = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "zc_test.so"))
zc.zc_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
for grid in (148, 148 * 4, 148 * 16):
    t = timeit(lambda: zc.zc_run(host.data_ptr(), dev_pack.data_ptr(), P, used // 4, pitch // 4, grid, st))
    print(f"zero-copy gather grid {grid}: {t:.3f} ms  payload {P*used/t/1e6:.1f} GB/s (rc {zc.zc_run(host.data_ptr(), dev_pack.data_ptr(), P, used // 4, pitch // 4, grid, st)})")
t = timeit(lambda: zc.zc_run(host.data_ptr(), dev_full.data_ptr(), P, pitch // 4, pitch // 4, 148 * 16, st))
print(f"zero-copy full read: {t:.3f} ms  {P*pitch/t/1e6:.1f} GB/s")
