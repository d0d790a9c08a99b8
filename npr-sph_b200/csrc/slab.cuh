// slab.cuh -- device-side pieces of the slab-decomposed (multi-GPU) step that both dist.cu and the
// fused force+integrate kernel of sph_passes.cu need: sentinel keys, the per-step counters, the
// classification of a freshly integrated particle (leaver / boundary layer / NaN / top layer).
#pragma once

#include "common.cuh"

namespace nprsph {

struct Migrant { float4 posid, velrho; };      // one particle changing rank (32 B)

// per-step counters, one block of CNT_WORDS per rank and exchanged with both neighbours
enum { CNT_LEAVE_L = 0, CNT_LEAVE_R, CNT_HALO_L, CNT_HALO_R, CNT_NAN, CNT_XMAX,
       // written by the host (reset_counts) for the neighbours' re-balancing decision:
       CNT_NOWN, CNT_FREE, CNT_WIDTH, CNT_CAPMIG, CNT_COST,
       CNT_WORDS = 12 };
// sticky error flags behind the three counter blocks (never cleared by the per-step memset)
// (+ the top occupied x layer of the last sort, written by the gather kernel of every prepare)
enum { ERR_IMMIGRANT = 0, ERR_OVERFLOW, STICKY_XMAX, ERR_WORDS = 8 };

// Sentinel keys of the slab sort.  All high bits set, so a sort on the low b bits keeps them behind
// every cell key (in this order) as long as the cell keys stay below 2^b - 4: the digit passes
// only have to cover the OCCUPIED part of the local grid (prepare_group), not the long empty
// stretch of box the last rank of a dam break owns.
constexpr uint32_t KEY_NAN = 0xFFFFFFFCu;       // stays on its rank for ever, in no cell
constexpr uint32_t KEY_GONE_L = 0xFFFFFFFDu;    // belongs to the left rank
constexpr uint32_t KEY_GONE_R = 0xFFFFFFFEu;    // belongs to the right rank
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;      // no particle (a lane without work)

// g: the local grid (own slab of W x layers plus R ghost layers on each side)
__device__ __forceinline__ uint32_t cell_key_slab(float x, float y, float z, const GridDev& g, int W, int R) {
    if (pos_is_nan(x, y, z)) return KEY_NAN;
    const int cxl = cell_x_unclamped(x, g);
    if (cxl < R) return KEY_GONE_L;
    if (cxl >= R + W) return KEY_GONE_R;
    const int cy = cell_coord(y, g.lo[1], g.inv_cell_d, g.dim[1]);
    const int cz = cell_coord(z, g.lo[2], g.inv_cell_d, g.dim[2]);
    return ((uint32_t)cxl * (uint32_t)g.dim[1] + (uint32_t)cy) * (uint32_t)g.dim[2] + (uint32_t)cz;
}

// What the kernels that produce next-step keys in slab mode need: the geometry the NEXT prepare
// will use (it differs from the current one in the step that re-balances the slab faces), the
// migration buffers and the counters.
struct SlabNext {
    GridDev g;                  // next local grid
    int W, R;                   // own x layers, reach
    Migrant* sendL;             // nullptr: leavers are only counted (scene distribution)
    Migrant* sendR;
    uint32_t cap_mig;
    uint32_t* counts;           // CNT_WORDS counters of this rank
    uint32_t* errs;             // ERR_WORDS sticky error flags
};

__device__ __forceinline__ void send_leaver(uint32_t key, const float4& p, const float4& v, const SlabNext& sn) {
    const bool goneL = key == KEY_GONE_L;
    const uint32_t slot = atomicAdd(sn.counts + (goneL ? CNT_LEAVE_L : CNT_LEAVE_R), 1u);
    Migrant* dst = goneL ? sn.sendL : sn.sendR;
    if (dst) {
        if (slot < sn.cap_mig) { dst[slot].posid = p; dst[slot].velrho = v; }
        else sn.errs[ERR_OVERFLOW] = 1u;
    }
}

// Boundary-layer / NaN counts for up to two freshly keyed own particles per thread.  EVERY lane of
// the warp calls it (KEY_NONE for a lane without a particle); warp-level only, no block barrier.
// (The top occupied x layer, which bounds the sort and the cell table of the last rank's long empty
// stretch of box, is no longer reduced here -- half a million warps on one address cost more than
// the rest of the kernel: the gather kernel reads it off the last sorted key, and a particle moves at
// most `reach` layers per step, see prepare_group.)
__device__ __forceinline__ void classify_counts(uint32_t ka, uint32_t kb, const SlabNext& sn) {
    const uint32_t plane = (uint32_t)sn.g.dim[1] * (uint32_t)sn.g.dim[2];
    const uint32_t xa = ka < KEY_NAN ? ka / plane : 0u, xb = kb < KEY_NAN ? kb / plane : 0u;
    const uint32_t twoR = (uint32_t)(2 * sn.R), W = (uint32_t)sn.W;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nL = __popc(__ballot_sync(0xffffffffu, ka < KEY_NAN && xa < twoR)) +
                        __popc(__ballot_sync(0xffffffffu, kb < KEY_NAN && xb < twoR));
    const uint32_t nR = __popc(__ballot_sync(0xffffffffu, ka < KEY_NAN && xa >= W)) +
                        __popc(__ballot_sync(0xffffffffu, kb < KEY_NAN && xb >= W));
    const uint32_t nN = __popc(__ballot_sync(0xffffffffu, ka == KEY_NAN)) +
                        __popc(__ballot_sync(0xffffffffu, kb == KEY_NAN));
    if (lane == 0) {
        if (nL) atomicAdd(sn.counts + CNT_HALO_L, nL);
        if (nR) atomicAdd(sn.counts + CNT_HALO_R, nR);
        if (nN) atomicAdd(sn.counts + CNT_NAN, nN);
    }
}

// One own particle's new key -> leaver into the migration buffer + the counts above.  EVERY lane of
// the warp calls it.
__device__ __forceinline__ void classify_key(uint32_t key, const float4& p, const float4& v, const SlabNext& sn) {
    if (key == KEY_GONE_L || key == KEY_GONE_R) send_leaver(key, p, v, sn);
    classify_counts(key, KEY_NONE, sn);
}

// The same for a thread on its own (the deferred-slot kernels: a few percent of the slots, of which
// only the boundary layers touch a counter).
__device__ __forceinline__ void classify_key_single(uint32_t key, const float4& p, const float4& v, const SlabNext& sn) {
    if (key == KEY_GONE_L || key == KEY_GONE_R) { send_leaver(key, p, v, sn); return; }
    if (key == KEY_NAN) { atomicAdd(sn.counts + CNT_NAN, 1u); return; }
    if (key == KEY_NONE) return;
    const uint32_t cxl = key / ((uint32_t)sn.g.dim[1] * (uint32_t)sn.g.dim[2]);
    if (cxl < (uint32_t)(2 * sn.R)) atomicAdd(sn.counts + CNT_HALO_L, 1u);
    if (cxl >= (uint32_t)sn.W) atomicAdd(sn.counts + CNT_HALO_R, 1u);
}

}  // namespace nprsph
