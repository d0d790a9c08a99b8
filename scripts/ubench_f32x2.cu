// Scratch micro-benchmark (B200): do the packed fp32x2 instructions (FADD2/FMUL2/FFMA2) save issue
// slots compared with scalar fp32?  Prints warp-instructions per clock per SM sub-partition and
// fp32 lane-operations per clock per SM for a few instruction mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
typedef unsigned long long u64;

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, float s, u64* cyc) {
    float a[8]; u64 p[8]; uint32_t m[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; m[i] = threadIdx.x + i; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
    u64 ps = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
    uint32_t ms = __float_as_uint(s);
    u64 t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));
            if (MODE == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));
            if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s), "f"(a[(i + 1) & 7]));
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
            if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
            if (MODE == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps), "l"(p[(i + 1) & 7]));
            if (MODE == 6) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));
                             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[i]) : "r"(ms), "r"(m[(i + 1) & 7])); }
            if (MODE == 7) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
                             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[i]) : "r"(ms), "r"(m[(i + 1) & 7])); }
            if (MODE == 8) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
                             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[i]) : "r"(ms), "r"(m[(i + 1) & 7]));
                             asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s)); }
            if (MODE == 9) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
                             asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s)); }
            if (MODE == 10) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[i]) : "r"(ms), "r"(m[(i + 1) & 7])); }
            if (MODE == 11) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
                              asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[(i + 4) & 7]) : "l"(ps)); }
            if (MODE == 12) { asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %2; @q add.rn.f32 %0, %0, %2; @q or.b32 %1, %1, %3;}"
                                           : "+f"(a[i]), "+r"(m[i]) : "f"(s), "r"(ms)); }
        }
    }
    u64 t1 = clock64();
    float r = 0; uint32_t x = 0;
    for (int i = 0; i < 8; i++) { r += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32)); x ^= m[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_slot, int lane_ops_per_slot, float* out, u64* cyc) {
    const int grid = 148;
    k<MODE><<<grid, 1024>>>(out, 1.0001f, cyc);
    k<MODE><<<grid, 1024>>>(out, 1.0001f, cyc);
    cudaDeviceSynchronize();
    u64 h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < grid; i++) c += (double)h[i]; c /= grid;
    const double warp_instr = 8.0 * ITER * 8 * instr_per_slot;       // per SMSP: 8 warps x ITER x 8 slots
    printf("%-34s cycles %9.0f  warp-instr/clk/SMSP %.3f  fp32 lane-ops/clk/SM %.1f\n", name, c, warp_instr / c,
           8.0 * ITER * 8 * lane_ops_per_slot * 32 * 4 / c);
}

int main() {
    float* out; u64* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("FADD", 1, 1, out, cyc);
    run<1>("FMUL", 1, 1, out, cyc);
    run<2>("FFMA 3-reg", 1, 1, out, cyc);
    run<3>("FADD2", 1, 2, out, cyc);
    run<4>("FMUL2", 1, 2, out, cyc);
    run<5>("FFMA2", 1, 2, out, cyc);
    run<10>("LOP3", 1, 0, out, cyc);
    run<6>("FADD + LOP3", 2, 1, out, cyc);
    run<7>("FADD2 + LOP3", 2, 2, out, cyc);
    run<8>("FADD2 + LOP3 + FADD", 3, 3, out, cyc);
    run<9>("FADD2 + FADD", 2, 3, out, cyc);
    run<11>("FMUL2 + FADD2", 2, 4, out, cyc);
    run<12>("FSETP + @p FADD + @p LOP", 3, 1, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
