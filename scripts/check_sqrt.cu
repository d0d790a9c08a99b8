// one-off: is the 5-instruction Markstein sequence bit-identical to sqrt.rn on every fp32 in [lo, hi]?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__global__ void k(uint32_t lo, uint32_t hi, unsigned long long* bad, uint32_t* first) {
    for (uint64_t b = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= hi; b += (uint64_t)gridDim.x * blockDim.x) {
        float x = __uint_as_float((uint32_t)b);
        float q = rsqrt_approx(x), r0 = x * q, e = fmaf(-r0, r0, x), r = fmaf(e, 0.5f * q, r0);
        if (__float_as_uint(r) != __float_as_uint(__fsqrt_rn(x))) { if (atomicAdd(bad, 1ull) == 0) *first = (uint32_t)b; }
    }
}
int main() {
    unsigned long long* bad; uint32_t* first; cudaMallocManaged(&bad, 8); cudaMallocManaged(&first, 4); *bad = 0; *first = 0;
    const float lo = 1e-30f, hi = 1e10f;
    k<<<148 * 16, 256>>>(*(uint32_t*)&lo, *(uint32_t*)&hi, bad, first);
    cudaDeviceSynchronize();
    printf("mismatches vs sqrt.rn over [%g, %g]: %llu (first bits 0x%08x) %s\n", lo, hi, *bad, *first, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
