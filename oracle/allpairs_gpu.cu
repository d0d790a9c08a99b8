/*
 * allpairs_gpu.cu -- all-pairs CUDA oracle.  TEST INFRASTRUCTURE ONLY (same rules as
 * sph_oracle.h: only tests/ may load liballpairs_gpu.so; the product never does).
 *
 * Literal transcription of the reference's O(N^2) loops (rho_pres_comp.glsl:46-54,
 * force_comp.glsl:48-62) with the canonical arithmetic of sph_oracle.c: thread i walks
 * j = 0..N-1 in order, every fp32 operation is individually rounded (built with
 * -fmad=false -prec-div=true -prec-sqrt=true -ftz=false), so results are bit-identical to the
 * C oracle.  It exists so that the 1M-particle configuration can be checked against a true
 * brute-force neighbour search in well under a minute.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "sph_oracle.h"

#define TILE 256

struct ap_consts {
    float h, mass315, den, mass, spiky, lap, gas_const, rest_rho, visc, g[3];
};

__global__ void __launch_bounds__(TILE)
ap_rho(const float4* __restrict__ rec, int n, ap_consts c, float2* __restrict__ rho_p,
       uint32_t* __restrict__ counts) {
    __shared__ float4 s_pos[TILE];
    const int i = blockIdx.x * TILE + threadIdx.x;
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) pi = rec[4 * (size_t)i];
    float rho = 0.0f;
    uint32_t cnt = 0;
    for (int base = 0; base < n; base += TILE) {
        const int j = base + threadIdx.x;
        if (j < n) s_pos[threadIdx.x] = rec[4 * (size_t)j];
        __syncthreads();
        const int lim = min(TILE, n - base);
        for (int t = 0; t < lim; t++) {
            const float4 pj = s_pos[t];
            const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            const float r = sqrtf((dx * dx + dy * dy) + dz * dz);
            if (r < c.h) {
                const float q = c.h * c.h - r * r;
                const float q3 = (q * q) * q;
                rho += (c.mass315 * q3) / c.den;
                cnt++;
            }
        }
        __syncthreads();
    }
    if (i < n) {
        const float pr = c.gas_const * (rho - c.rest_rho);
        rho_p[i] = make_float2(rho, (pr < 0.0f) ? 0.0f : pr);
        if (counts) counts[i] = cnt;
    }
}

__global__ void __launch_bounds__(TILE)
ap_force(const float4* __restrict__ rec, int n, ap_consts c, float4* __restrict__ f_out,
         uint32_t* __restrict__ counts) {
    __shared__ float4 s_pos[TILE], s_vel[TILE];
    __shared__ float2 s_rp[TILE];
    const int i = blockIdx.x * TILE + threadIdx.x;
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), vi = pi, ei = pi;
    if (i < n) { pi = rec[4 * (size_t)i]; vi = rec[4 * (size_t)i + 1]; ei = rec[4 * (size_t)i + 3]; }
    float px = 0.f, py = 0.f, pz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
    uint32_t cnt = 0;
    for (int base = 0; base < n; base += TILE) {
        const int j = base + threadIdx.x;
        if (j < n) {
            s_pos[threadIdx.x] = rec[4 * (size_t)j];
            s_vel[threadIdx.x] = rec[4 * (size_t)j + 1];
            const float4 e = rec[4 * (size_t)j + 3];
            s_rp[threadIdx.x] = make_float2(e.x, e.y);
        }
        __syncthreads();
        const int lim = min(TILE, n - base);
        for (int t = 0; t < lim; t++) {
            if (base + t == i) continue;
            const float4 pj = s_pos[t];
            const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            const float r = sqrtf((dx * dx + dy * dy) + dz * dz);
            if (r < c.h) {
                const float4 vj = s_vel[t];
                const float2 rp = s_rp[t];
                const float hr = c.h - r;
                const float s = (((c.mass * (ei.y + rp.y)) / (2.0f * rp.x)) * c.spiky) * (hr * hr);
                px -= s * (dx / r); py -= s * (dy / r); pz -= s * (dz / r);
                vx += (((c.mass * (vj.x - vi.x)) / rp.x) * c.lap) * hr;
                vy += (((c.mass * (vj.y - vi.y)) / rp.x) * c.lap) * hr;
                vz += (((c.mass * (vj.z - vi.z)) / rp.x) * c.lap) * hr;
                cnt++;
            }
        }
        __syncthreads();
    }
    if (i < n) {
        float4 f;
        f.x = (px + vx * c.visc) + ei.x * c.g[0];
        f.y = (py + vy * c.visc) + ei.x * c.g[1];
        f.z = (pz + vz * c.visc) + ei.x * c.g[2];
        f.w = 0.f;
        f_out[i] = f;
        if (counts) counts[i] = cnt;
    }
}

/* Conditioning scale of the force sums (see oracle_force_scale in sph_oracle.c). */
__global__ void __launch_bounds__(TILE)
ap_force_scale(const float4* __restrict__ rec, int n, ap_consts c, float4* __restrict__ out) {
    __shared__ float4 s_pos[TILE], s_vel[TILE];
    __shared__ float2 s_rp[TILE];
    const int i = blockIdx.x * TILE + threadIdx.x;
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), vi = pi, ei = pi;
    if (i < n) { pi = rec[4 * (size_t)i]; vi = rec[4 * (size_t)i + 1]; ei = rec[4 * (size_t)i + 3]; }
    double ax = 0, ay = 0, az = 0;
    for (int base = 0; base < n; base += TILE) {
        const int j = base + threadIdx.x;
        if (j < n) {
            s_pos[threadIdx.x] = rec[4 * (size_t)j];
            s_vel[threadIdx.x] = rec[4 * (size_t)j + 1];
            const float4 e = rec[4 * (size_t)j + 3];
            s_rp[threadIdx.x] = make_float2(e.x, e.y);
        }
        __syncthreads();
        const int lim = min(TILE, n - base);
        for (int t = 0; t < lim; t++) {
            if (base + t == i) continue;
            const float4 pj = s_pos[t];
            const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            const float r = sqrtf((dx * dx + dy * dy) + dz * dz);
            if (r < c.h) {
                const float4 vj = s_vel[t];
                const float2 rp = s_rp[t];
                const float hr = c.h - r;
                const float s = (((c.mass * (ei.y + rp.y)) / (2.0f * rp.x)) * c.spiky) * (hr * hr);
                ax += fabs((double)(s * (dx / r))) + fabs((double)((((c.mass * (vj.x - vi.x)) / rp.x) * c.lap) * hr) * c.visc);
                ay += fabs((double)(s * (dy / r))) + fabs((double)((((c.mass * (vj.y - vi.y)) / rp.x) * c.lap) * hr) * c.visc);
                az += fabs((double)(s * (dz / r))) + fabs((double)((((c.mass * (vj.z - vi.z)) / rp.x) * c.lap) * hr) * c.visc);
            }
        }
        __syncthreads();
    }
    if (i < n)
        out[i] = make_float4((float)(ax + fabs((double)ei.x * c.g[0])), (float)(ay + fabs((double)ei.x * c.g[1])),
                             (float)(az + fabs((double)ei.x * c.g[2])), 0.f);
}

static ap_consts make_consts(const oracle_params* p) {
    ap_consts c;
    c.h = p->smoothing_coeff * p->particle_radius;
    c.mass315 = p->mass * 315.0f;
    c.den = (64.0f * p->pi) * (float)pow((double)c.h, 9.0);
    const float h6 = (float)pow((double)c.h, 6.0);
    c.mass = p->mass;
    c.spiky = -45.0f / (p->pi * h6);
    c.lap = 45.0f / (p->pi * h6);
    c.gas_const = p->gas_const;
    c.rest_rho = p->resting_rho;
    c.visc = p->visc;
    for (int a = 0; a < 3; a++) c.g[a] = p->gravity[a];
    return c;
}

#define APCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "allpairs_gpu: %s: %s\n", #x, cudaGetErrorString(e_)); rc = -1; goto done; } } while (0)

/* which: 0 = rho pass (writes extras[0..1]), 1 = force pass (writes force.xyz),
 * 2 = force conditioning scale (particles untouched; `counts` then points to float[3n]).
 * particles: host array of n records, updated in place.  counts: nullable host array. */
extern "C" int oracle_gpu_pass(int which, float* particles, int n, const oracle_params* p,
                               uint32_t* counts) {
    int rc = 0;
    float4* d_rec = nullptr; void* d_out = nullptr; uint32_t* d_cnt = nullptr;
    const ap_consts c = make_consts(p);
    const int blocks = (n + TILE - 1) / TILE;
    float* h_out = nullptr;
    if (n <= 0) return 0;
    APCK(cudaMalloc(&d_rec, (size_t)n * 64));
    APCK(cudaMalloc(&d_out, (size_t)n * 16));
    APCK(cudaMalloc(&d_cnt, (size_t)n * 4));
    APCK(cudaMemcpy(d_rec, particles, (size_t)n * 64, cudaMemcpyHostToDevice));
    if (which == 0)      ap_rho<<<blocks, TILE>>>(d_rec, n, c, (float2*)d_out, d_cnt);
    else if (which == 1) ap_force<<<blocks, TILE>>>(d_rec, n, c, (float4*)d_out, d_cnt);
    else                 ap_force_scale<<<blocks, TILE>>>(d_rec, n, c, (float4*)d_out);
    APCK(cudaGetLastError());
    APCK(cudaDeviceSynchronize());
    h_out = (float*)malloc((size_t)n * 16);
    APCK(cudaMemcpy(h_out, d_out, (size_t)n * (which == 0 ? 8 : 16), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
        float* r = particles + (size_t)i * ORACLE_REC;
        if (which == 0) { r[12] = h_out[2 * i]; r[13] = h_out[2 * i + 1]; }
        else if (which == 1) { r[8] = h_out[4 * i]; r[9] = h_out[4 * i + 1]; r[10] = h_out[4 * i + 2]; }
        else { float* o = (float*)counts; o[3 * (size_t)i] = h_out[4 * i]; o[3 * (size_t)i + 1] = h_out[4 * i + 1]; o[3 * (size_t)i + 2] = h_out[4 * i + 2]; }
    }
    if (counts && which != 2) APCK(cudaMemcpy(counts, d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost));
done:
    free(h_out);
    cudaFree(d_rec); cudaFree(d_out); cudaFree(d_cnt);
    return rc;
}
