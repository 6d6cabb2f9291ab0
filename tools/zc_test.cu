// Developer experiment: gather the used SH coefficients straight from pinned host memory (zero-copy over PCIe).
#include <cuda_runtime.h>
#include <stdint.h>
extern "C" __global__ void zc_gather(const float* __restrict__ src, float* __restrict__ dst, int P, int used, int pitch) {
    const long long n = (long long)P * used;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / used; const int k = (int)(i - p * used);
        dst[i] = __ldcs(src + p * pitch + k);
    }
}
extern "C" int zc_run(const float* host, float* dst, int P, int used, int pitch, int grid, void* stream) {
    zc_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(host, dst, P, used, pitch);
    return (int)cudaGetLastError();
}
