// common.cuh -- device-side parameter blocks and helpers shared by all kernels.
// sm_100a only; there is deliberately no host/CPU implementation of any pass.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace nprsph {

// Column records handed from the density pass to the force pass (sph_passes.cu "column records"):
// one 64-bit record per non-empty column of a walk, two record planes per slot pair, one control
// word per slot pair.  A walk visits at most (2*reach+1)^2 columns; the buffer is laid out for the
// reach of the current grid (rec_cols_of(g.reach)), so the reference's own default (h = 4 lattice
// spacings, cell = h/4, reach 4: 81 columns of <= 12 candidates) replays records like the dam break
// (h = 2 spacings, reach 2: 25 columns).
constexpr int REC_REACH_MAX = 4;                                       // reach the records support
__host__ __device__ constexpr uint32_t rec_cols_of(int reach) { return (uint32_t)((2 * reach + 1) * (2 * reach + 1)); }
// Behind the records and the control words: the queue of deferred slots (sph_passes.cu "deferred
// singles"): two counters (density pass, force pass), then up to n slot numbers.
__host__ __device__ constexpr size_t rec_queue_offset(size_t n, uint32_t cols) { return ((4 * (size_t)cols + 1) * ((n + 1) / 2) + 3) & ~(size_t)3; }
// words of the record buffer for a capacity of n slots
constexpr size_t rec_buffer_words(size_t n, uint32_t cols) { return rec_queue_offset(n, cols) + 4 + n; }

// Uniform-grid definition (DESIGN.md "Grid").  The same arithmetic is restated in
// oracle/sph_oracle.c:oracle_grid_setup / cell_key so keys can be compared bit-exactly.
struct GridDev {
    float lo[3];
    float inv_cell;
    int   dim[3];
    int   reach;          // cells walked on each side of the home cell
    uint32_t num_cells;   // sentinel key (NaN positions) == num_cells
    // Slab decomposition (multi-GPU): dim[0] is the LOCAL number of x cells (own slab plus `reach`
    // ghost layers on each side), x_off the global x index of local cell 0 and dimx_global the
    // global number of x cells.  Single GPU: x_off = 0, dimx_global = dim[0].
    int   x_off;
    int   dimx_global;
    int   pad0;           // (keeps the struct free of implicit padding: it is compared with memcmp)
    double inv_cell_d;    // 1 / cell in double: cell coordinates are computed in fp64 (cell_coord)
};

// Run-time constants of the three passes.  Sources: ConstantsUniform / BoundaryUniform
// (Main.cpp:110-122) and the shader consts (rho_pres_comp.glsl:5,8,33; force_comp.glsl:33;
// integrate_comp.glsl:8,33).
struct SphDev {
    float h;            // smoothing_coeff * particle_radius   (rho_pres_comp.glsl:40)
    float h2;           // h*h
    float r2_max;       // smallest t with sqrt_rn(t) >= h  =>  (length(d) < h) == (r2 < r2_max)
    float cull2;        // (h in cell units + rounding margin)^2: column-footprint cull threshold
    float rho_coef;     // mass*315 / (64*pi*h^9)             (rho_pres_comp.glsl:52)
    float pres_coef;    // -mass*spiky/2 = mass*45/(2*pi*h^6) (force_comp.glsl:41,59)
    float visc_coef;    // visc*mass*laplacian                (force_comp.glsl:42,60,63)
    float gas_const;
    float rest_rho;
    float g[3];
    float damping;
    float dt;
    float lower[3];
    float upper[3];
    float zero;         // 0.0f the compiler cannot see (VecConsts)
    float one;          // 1.0f the compiler cannot see: fma(x, one, y) is an exactly rounded add that
                        // ptxas cannot contract with a preceding multiply (it fuses mul.rn.f32x2 +
                        // add.rn.f32x2 into FFMA2 even with explicit rounding modifiers)
};

__device__ __forceinline__ int cell_coord(float x, float lo, double inv_cell, int dim) {
    // (x - lo) * inv_cell in fp64, two individually rounded operations: the error of the cell
    // coordinate (~ dim * 2^-52 cells) is far below the 2^-14 widening of the cell, so `reach` cells
    // reach every neighbour whatever the size of the grid, and the cell size does not depend on it
    double u = __dmul_rn(__dsub_rn((double)x, (double)lo), inv_cell);
    if (!(u >= 0.0)) u = 0.0;
    const double top = (double)(dim - 1);
    if (u > top) u = top;
    return (int)u;
}

__device__ __forceinline__ bool pos_is_nan(float x, float y, float z) {
    return (x != x) || (y != y) || (z != z);
}

// local x cell index; the global index is computed first so that every rank assigns a particle
// to the same global cell bit for bit.  May fall outside [0, dim[0]) for a particle that left
// the slab.
__device__ __forceinline__ int cell_x_unclamped(float x, const GridDev& g) {
    return cell_coord(x, g.lo[0], g.inv_cell_d, g.dimx_global) - g.x_off;
}
__device__ __forceinline__ int cell_x(float x, const GridDev& g) {
    return min(max(cell_x_unclamped(x, g), 0), g.dim[0] - 1);
}

__device__ __forceinline__ uint32_t cell_key(float x, float y, float z, const GridDev& g) {
    if (pos_is_nan(x, y, z)) return g.num_cells;
    const int cx = cell_x(x, g);
    const int cy = cell_coord(y, g.lo[1], g.inv_cell_d, g.dim[1]);
    const int cz = cell_coord(z, g.lo[2], g.inv_cell_d, g.dim[2]);
    return ((uint32_t)cx * (uint32_t)g.dim[1] + (uint32_t)cy) * (uint32_t)g.dim[2] + (uint32_t)cz;
}

// The neighbour predicate of the shaders, `length(pos_i - pos_j) < h`
// (rho_pres_comp.glsl:48-50, force_comp.glsl:55-57), evaluated as r2 < r2_max with r2 built
// from individually rounded operations in the oracle's order: (dx*dx + dy*dy) + dz*dz.
__device__ __forceinline__ float dist2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// GLSL max(a,b) = (a < b) ? b : a
__device__ __forceinline__ float glsl_max(float a, float b) { return (a < b) ? b : a; }

__device__ __forceinline__ float eos_pressure(float rho, const SphDev& s) {
    // max(GAS_CONST * (rho - resting_rho), 0)   rho_pres_comp.glsl:58
    return glsl_max(__fmul_rn(s.gas_const, __fsub_rn(rho, s.rest_rho)), 0.0f);
}


// ---- empty runs of the cell table -------------------------------------------------------------------------
// A thread of the table builders fills short empty runs itself and queues the others for
// k_fill_gaps (grid.cu), one thread block per queue entry.  A long run (the empty stretch of box
// ahead of a dam break, the empty upper half of every x plane of a deep scene) is cut into pieces of
// GAP_PIECE cells, so no block ever writes more than 256 KB.  (Round 1 kept the "huge" runs in a
// separate 64-entry list filled by the whole grid and let an overflow fall back to ONE block per run:
// a 512-cell-deep scene has a huge run in every x plane, and the first NaN particle then sent its
// 135-million-cell run to a single block -- 11.9 ms per step instead of 0.45.)
// ctl words: [0] queue length; [4], [5] tail state of the single-context table (grid.cu).
constexpr uint32_t GAP_INLINE = 64;          // runs up to this length are written by the finder
constexpr uint32_t GAP_PIECE = 1u << 16;     // queue entries hold at most this many cells
constexpr size_t GAP_CTL_BYTES = 32;
__device__ __forceinline__ void push_gap(uint32_t lo, uint32_t len, uint32_t slot,
                                         uint4* __restrict__ gap_list, uint32_t* __restrict__ ctl) {
    const uint32_t pieces = (len + GAP_PIECE - 1u) / GAP_PIECE;
    uint32_t k = atomicAdd(ctl, pieces);
    for (uint32_t o = 0; o < len; o += GAP_PIECE, ++k)
        gap_list[k] = make_uint4(lo + o, min(GAP_PIECE, len - o), slot, 0u);
}

// ---- pass 3 for one particle ----------------------------------------------------------------------------
// Static colliders (README.md:59 "Add objects for particles to collide with"):
// not in the reference's shaders; the response mirrors its wall rule (integrate_comp.glsl:46-77:
// put the particle on the surface, multiply the normal velocity by -DAMPING).  Kernel parameter of
// the integrate kernels only.  a.w = kind (0 sphere: a.xyz centre, b.x radius; 1 box: a.xyz lower
// corner, b.xyz upper corner).  Restated op for op in oracle/sph_oracle.c:collide_one.
constexpr int MAX_COLLIDERS = 8;
struct ColliderSet {
    int n;
    int pad[3];
    float4 a[MAX_COLLIDERS];
    float4 b[MAX_COLLIDERS];
};

__device__ __forceinline__ void collide_sphere(float4& p, float4& v, float cx, float cy, float cz,
                                               float R, float damping) {
    const float dx = __fsub_rn(p.x, cx), dy = __fsub_rn(p.y, cy), dz = __fsub_rn(p.z, cz);
    const float r2 = dist2_exact(dx, dy, dz);
    if (!(r2 < __fmul_rn(R, R))) return;
    const float r = __fsqrt_rn(r2);
    float nx = 0.0f, ny = 1.0f, nz = 0.0f;                   // a particle exactly at the centre leaves upwards
    if (r > 0.0f) { nx = __fdiv_rn(dx, r); ny = __fdiv_rn(dy, r); nz = __fdiv_rn(dz, r); }
    p.x = __fadd_rn(cx, __fmul_rn(R, nx));
    p.y = __fadd_rn(cy, __fmul_rn(R, ny));
    p.z = __fadd_rn(cz, __fmul_rn(R, nz));
    const float vn = __fadd_rn(__fadd_rn(__fmul_rn(v.x, nx), __fmul_rn(v.y, ny)), __fmul_rn(v.z, nz));
    const float k = __fmul_rn(__fadd_rn(1.0f, damping), vn);    // v_n -> -damping * v_n
    v.x = __fsub_rn(v.x, __fmul_rn(k, nx));
    v.y = __fsub_rn(v.y, __fmul_rn(k, ny));
    v.z = __fsub_rn(v.z, __fmul_rn(k, nz));
}

__device__ __forceinline__ void collide_box(float4& p, float4& v, const float4& lo, const float4& hi,
                                            float damping) {
    if (!(p.x > lo.x && p.x < hi.x && p.y > lo.y && p.y < hi.y && p.z > lo.z && p.z < hi.z)) return;
    // nearest face, first minimum in the order -x, +x, -y, +y, -z, +z
    const float pen[6] = {__fsub_rn(p.x, lo.x), __fsub_rn(hi.x, p.x), __fsub_rn(p.y, lo.y),
                          __fsub_rn(hi.y, p.y), __fsub_rn(p.z, lo.z), __fsub_rn(hi.z, p.z)};
    int best = 0;
    float bp = pen[0];
#pragma unroll
    for (int f = 1; f < 6; f++) if (pen[f] < bp) { bp = pen[f]; best = f; }
    const float nd = -damping;
    if (best == 0)      { p.x = lo.x; v.x = __fmul_rn(v.x, nd); }
    else if (best == 1) { p.x = hi.x; v.x = __fmul_rn(v.x, nd); }
    else if (best == 2) { p.y = lo.y; v.y = __fmul_rn(v.y, nd); }
    else if (best == 3) { p.y = hi.y; v.y = __fmul_rn(v.y, nd); }
    else if (best == 4) { p.z = lo.z; v.z = __fmul_rn(v.z, nd); }
    else                { p.z = hi.z; v.z = __fmul_rn(v.z, nd); }
}

// integrate_comp.glsl:35-82 for one particle: a = F/rho, v += dt a, x += dt v (:41-43), then the
// colliders (extension), then the box walls (:46-77).  The shader's arithmetic operation by
// operation (no contraction), so given identical inputs the result is bit-identical to the oracle's.
__device__ __forceinline__ void integrate_particle(float4& p, float4& v, const float4& f,
                                                   const SphDev& sp, const ColliderSet& cs) {
    const float rho = v.w, nd = -sp.damping;
    v.x = __fadd_rn(v.x, __fmul_rn(sp.dt, __fdiv_rn(f.x, rho)));
    v.y = __fadd_rn(v.y, __fmul_rn(sp.dt, __fdiv_rn(f.y, rho)));
    v.z = __fadd_rn(v.z, __fmul_rn(sp.dt, __fdiv_rn(f.z, rho)));
    p.x = __fadd_rn(p.x, __fmul_rn(sp.dt, v.x));
    p.y = __fadd_rn(p.y, __fmul_rn(sp.dt, v.y));
    p.z = __fadd_rn(p.z, __fmul_rn(sp.dt, v.z));
    for (int c = 0; c < cs.n; c++) {
        const float4 a = cs.a[c], b = cs.b[c];
        if (a.w == 0.0f) collide_sphere(p, v, a.x, a.y, a.z, b.x, sp.damping);
        else             collide_box(p, v, a, b, sp.damping);
    }
    if (p.x < sp.lower[0])      { p.x = sp.lower[0]; v.x = __fmul_rn(v.x, nd); }
    else if (p.x > sp.upper[0]) { p.x = sp.upper[0]; v.x = __fmul_rn(v.x, nd); }
    if (p.y < sp.lower[1])      { p.y = sp.lower[1]; v.y = __fmul_rn(v.y, nd); }
    else if (p.y > sp.upper[1]) { p.y = sp.upper[1]; v.y = __fmul_rn(v.y, nd); }
    if (p.z < sp.lower[2])      { p.z = sp.lower[2]; v.z = __fmul_rn(v.z, nd); }
    else if (p.z > sp.upper[2]) { p.z = sp.upper[2]; v.z = __fmul_rn(v.z, nd); }
}

}  // namespace nprsph
