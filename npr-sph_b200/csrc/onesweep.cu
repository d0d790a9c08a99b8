// onesweep.cu -- hand-written least-significant-digit radix sort of (cell key, slot) pairs.
//
// One histogram kernel reads the keys once and counts every digit position; then one "onesweep"
// kernel per digit moves the pairs: each CTA takes a tile of 4096 pairs (dynamic tile ticket, so
// every predecessor tile is already resident), ranks its keys stably inside each warp,
// publishes its per-digit counts to a tile-status array and resolves its global offsets by
// decoupled look-back over the predecessors' status words (count and flag share one 32-bit word,
// so no fence is needed).  Pairs are staged through shared memory so the global scatter is
// written in digit-contiguous runs.
//
// Only the low `key_bits` bits are sorted (ceil(log2(num_cells+1)) for cell keys).  Digits are
// 8 bits wide, or 9 bits when that saves a whole pass (27-bit keys of the 16 Mi-particle grid:
// 3 passes instead of 4).
// No reference counterpart: the reference searches neighbours all-pairs
// (rho_pres_comp.glsl:46, force_comp.glsl:48).
#include "sort.cuh"

#ifndef NPRSPH_SORT_BALLOT
#define NPRSPH_SORT_BALLOT 1
#endif

namespace nprsph {

namespace {

constexpr uint32_t FLAG_AGG  = 1u << 30;   // tile count published
constexpr uint32_t FLAG_INCL = 1u << 31;   // inclusive prefix published
constexpr uint32_t FLAG_ANY  = FLAG_AGG | FLAG_INCL;
constexpr uint32_t VALUE_MASK = FLAG_AGG - 1;

struct DigitPlan {                 // digit position of every pass
    int passes;
    int shift[SORT_MAX_PASSES];
    uint32_t mask[SORT_MAX_PASSES];
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---- digit histograms for all passes in one read of the keys --------------------------------
// Each thread walks a contiguous run of keys and run-length-compresses equal digits before
// touching shared memory: cell-ordered keys share their high digits over long runs, which
// would otherwise serialise the shared-memory atomics 32 ways.
constexpr int HIST_THREADS = 256;
constexpr int HIST_ITEMS = 16;

__global__ void __launch_bounds__(HIST_THREADS)
k_radix_hist(const uint32_t* __restrict__ keys, uint32_t n, DigitPlan plan, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[SORT_MAX_PASSES][SORT_MAX_RADIX];
    for (int i = threadIdx.x; i < SORT_MAX_PASSES * SORT_MAX_RADIX; i += HIST_THREADS) (&sh[0][0])[i] = 0;
    __syncthreads();

    const uint64_t chunk = (uint64_t)HIST_THREADS * HIST_ITEMS;
    for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < n; base += (uint64_t)gridDim.x * chunk) {
        const uint64_t first = base + (uint64_t)threadIdx.x * HIST_ITEMS;
        uint32_t k[HIST_ITEMS];
        int cnt = 0;
        if (first + HIST_ITEMS <= n) {
            const uint4* p = reinterpret_cast<const uint4*>(keys + first);   // first % 16 == 0
#pragma unroll
            for (int v = 0; v < HIST_ITEMS / 4; v++) {
                uint4 q = __ldg(p + v);
                k[4 * v] = q.x; k[4 * v + 1] = q.y; k[4 * v + 2] = q.z; k[4 * v + 3] = q.w;
            }
            cnt = HIST_ITEMS;
        } else {
#pragma unroll
            for (int v = 0; v < HIST_ITEMS; v++)
                if (first + v < n) { k[v] = keys[first + v]; cnt = v + 1; }
        }
        for (int p = 0; p < plan.passes; p++) {
            const int shift = plan.shift[p];
            const uint32_t mask = plan.mask[p];
            uint32_t run_d = cnt ? ((k[0] >> shift) & mask) : 0xFFFFFFFFu;
            uint32_t run = cnt ? 1u : 0u;
#pragma unroll
            for (int v = 1; v < HIST_ITEMS; v++) {
                if (v < cnt) {
                    const uint32_t d = (k[v] >> shift) & mask;
                    if (d != run_d) { atomicAdd(&sh[p][run_d], run); run_d = d; run = 0; }
                    run++;
                }
            }
            // the pending run: one atomic per warp when the whole warp ended on the same digit
            // (the usual case for the high digits of cell-ordered keys), else one per thread
            const uint32_t d0 = __shfl_sync(0xffffffffu, run_d, 0);
            if (__all_sync(0xffffffffu, run_d == d0)) {
                const uint32_t total = __reduce_add_sync(0xffffffffu, run);
                if ((threadIdx.x & 31) == 0 && total) atomicAdd(&sh[p][d0], total);
            } else if (run) {
                atomicAdd(&sh[p][run_d], run);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * SORT_MAX_RADIX; i += HIST_THREADS) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

// exclusive scan of one value per thread over the CTA (SORT_WARPS warps)
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp_tot /*[SORT_WARPS]*/) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++)
        if ((uint32_t)w < warp) off += s_warp_tot[w];
    __syncthreads();
    return off + incl - v;
}

// ---- one digit pass ---------------------------------------------------------------------------
// BITS: digit width the kernel is built for (radix 2^BITS, DPT = radix/256 digits per thread in the
// scan / look-back phase); the pass may use fewer bits (mask).
template <int BITS, bool IOTA>
__global__ void __launch_bounds__(SORT_THREADS)
k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
           uint32_t mask, const uint32_t* __restrict__ hist, uint32_t* __restrict__ tile_counter,
           volatile uint32_t* __restrict__ status) {
    constexpr int RDX = 1 << BITS;
    constexpr int DPT = RDX >= SORT_THREADS ? RDX / SORT_THREADS : 1;     // digits a thread owns in the scan / look-back
    static_assert(SORT_WARPS * RDX <= 2 * SORT_TILE, "warp histograms must fit the staging area");
    // [0, TILE): staged keys.  [TILE, 2*TILE): the per-warp digit counters while ranking, then the
    // staged values (every read of the counters is finished before the first value is staged).
    __shared__ uint32_t s_buf[2 * SORT_TILE];
    __shared__ uint32_t s_out_base[RDX];      // global position of local slot 0 of each digit run
    __shared__ uint32_t s_local_off[RDX];
    __shared__ uint32_t s_scan[SORT_WARPS];
    __shared__ uint32_t s_tile;
    uint32_t* s_keys = s_buf;
    uint32_t* s_vals = s_buf + SORT_TILE;
    // [SORT_WARPS][RDX]; with 16 warps and 512 digits it takes both halves (the keys are staged after
    // the last read of the counters, too)
    uint32_t* s_warp_hist = s_buf + (SORT_WARPS * RDX <= SORT_TILE ? SORT_TILE : 0);
    const bool owner = (int)threadIdx.x * DPT < RDX;        // 512 threads, 256 digits: the upper half owns none

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < SORT_WARPS * RDX; i += SORT_THREADS) s_warp_hist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * SORT_TILE;
    const uint32_t warp_base = tile_base + warp * (SORT_ITEMS * 32);

    // warp-striped load: item k of lane l sits at warp_base + k*32 + l (order = warp, k, lane)
    uint32_t key[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t idx = warp_base + k * 32 + lane;
        key[k] = (idx < n) ? keys_in[idx] : 0xFFFFFFFFu;
    }

    // Stable rank of every key among the keys of its warp with the same digit.  The votes of all
    // rows are issued first (independent); then one shared-memory atomic per distinct digit per
    // row (by the lowest peer lane) claims the rank base, broadcast back with a shuffle.  Atomics
    // of one warp to one address retire in program order, so row k ranks below row k+1: the sort
    // stays stable, and the 16 rows overlap instead of forming one load -> store chain.
    uint32_t rank[SORT_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* my_hist = s_warp_hist + warp * RDX;
    uint32_t peers[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        // MATCH.ANY costs one round per distinct value in the warp.  Cell-ordered keys give rows
        // whose digits are either all equal or count upwards (consecutive cells; after the first
        // pass, keys one radix apart), i.e. all distinct: both cases are recognised with two
        // shuffles and a vote, and only irregular rows pay for the match.
        const uint32_t d = (key[k] >> shift) & mask;
        const uint32_t e = (d - __shfl_sync(0xffffffffu, d, 0)) & mask;     // offset from lane 0, cyclic
        const uint32_t below_e = __shfl_up_sync(0xffffffffu, e, 1);
        const bool rising = lane == 0 || e > below_e;
        if (!__any_sync(0xffffffffu, e != 0u))      peers[k] = 0xffffffffu;           // one digit
        else if (__all_sync(0xffffffffu, rising))   peers[k] = 1u << lane;            // all distinct
        else {
#if NPRSPH_SORT_BALLOT == 1
            // Irregular row (in a flowing fluid ~10 % of the particles change cell per step, so almost
            // every row holds one): peers from one vote per digit bit -- a fixed 9 votes -- instead of
            // MATCH.ANY's one round per distinct value (20-30 in a cell-ordered row).  Measured on B200,
            // 16 Mi evolved dam break: 0.63 -> 0.56 ms per sort; voting only on the bits in which the
            // row differs (a REDUX + a data-dependent loop) gave the gain back (0.62 ms).
            uint32_t pm = 0xffffffffu;
#pragma unroll
            for (int bit = 0; bit < BITS; bit++) {
                const bool on = (d >> bit) & 1u;
                const uint32_t v = __ballot_sync(0xffffffffu, on);
                pm &= on ? v : ~v;
            }
            peers[k] = pm;
#else
            peers[k] = __match_any_sync(0xffffffffu, d);
#endif
        }
    }
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t d = (key[k] >> shift) & mask;
        const uint32_t below = peers[k] & lt_mask;
        uint32_t base = 0;
        if (below == 0) base = atomicAdd(&my_hist[d], (uint32_t)__popc(peers[k]));
        base = __shfl_sync(0xffffffffu, base, __ffs(peers[k]) - 1);
        rank[k] = base + __popc(below);
    }
    __syncthreads();

    // thread t owns digits t*DPT .. t*DPT+DPT-1: counts across warps -> exclusive warp offsets + tile count
    uint32_t tile_count[DPT], hist_own[DPT];
    uint32_t sum_tc = 0, sum_h = 0;
#pragma unroll
    for (int q = 0; q < DPT; q++) {
        const uint32_t d_own = tid * DPT + q;
        uint32_t tc = 0;
        tile_count[q] = 0; hist_own[q] = 0;
        if (owner) {
#pragma unroll
            for (int w = 0; w < SORT_WARPS; w++) {
                const uint32_t c = s_warp_hist[w * RDX + d_own];
                s_warp_hist[w * RDX + d_own] = tc;
                tc += c;
            }
            tile_count[q] = tc;
            status[(uint64_t)tile * RDX + d_own] = tc | (tile == 0 ? FLAG_INCL : FLAG_AGG);
            hist_own[q] = hist[d_own];
        }
        sum_tc += tc; sum_h += hist_own[q];
    }
    uint32_t global_excl = block_excl_scan_256(sum_h, s_scan);
    uint32_t local_off = block_excl_scan_256(sum_tc, s_scan);

    // decoupled look-back over predecessor tiles, the DPT digits of a thread interleaved
    // Each step reads a window of LOOK predecessors at once: the loads are independent, so a walk
    // of depth k costs ceil(k / LOOK) L2 round trips instead of k (the walk was a third of the
    // kernel's stall samples when it read one status word at a time).
    constexpr int LOOK = 8;
    uint32_t excl[DPT];
    if (tile > 0 && owner) {
#pragma unroll
        for (int q = 0; q < DPT; q++) {
            const uint32_t d_own = tid * DPT + q;
            uint32_t acc = 0;
            int64_t t = (int64_t)tile - 1;               // next predecessor to consume
            bool done = false;
            while (!done) {
                uint32_t v[LOOK];
#pragma unroll
                for (int w = 0; w < LOOK; w++)
                    v[w] = (t - w >= 0) ? status[(uint64_t)(t - w) * RDX + d_own] : (uint32_t)(1u << 31);   // before tile 0: inclusive prefix 0
#pragma unroll
                for (int w = 0; w < LOOK; w++) {
                    if (done) break;
                    if ((v[w] & FLAG_ANY) == 0) break;   // not published yet: poll again from here
                    acc += v[w] & VALUE_MASK;
                    t--;
                    if (v[w] & FLAG_INCL) done = true;
                }
            }
            excl[q] = acc;
            status[(uint64_t)tile * RDX + d_own] = ((acc + tile_count[q]) & VALUE_MASK) | FLAG_INCL;
        }
    } else {
#pragma unroll
        for (int q = 0; q < DPT; q++) excl[q] = 0;
    }
#pragma unroll
    for (int q = 0; q < DPT; q++) {
        if (!owner) break;
        const uint32_t d_own = tid * DPT + q;
        s_out_base[d_own] = global_excl + excl[q] - local_off;
        s_local_off[d_own] = local_off;
        global_excl += hist_own[q];
        local_off += tile_count[q];
    }
    __syncthreads();

    // final position of every pair inside the tile (last reads of the warp counters) ...
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t d = (key[k] >> shift) & mask;
        rank[k] += s_local_off[d] + my_hist[d];
    }
    __syncthreads();
    // ... then stage the pairs in tile-sorted order (values overwrite the counters)
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t idx = warp_base + k * 32 + lane;
        s_keys[rank[k]] = key[k];
        s_vals[rank[k]] = IOTA ? idx : ((idx < n) ? vals_in[idx] : 0u);
    }
    __syncthreads();

    // digit-contiguous scatter; padding keys (last tile) sort to the tail of the tile
    const uint32_t valid = (n - tile_base < (uint32_t)SORT_TILE) ? (n - tile_base) : (uint32_t)SORT_TILE;
#pragma unroll
    for (int m = 0; m < SORT_ITEMS; m++) {
        const uint32_t s = tid + m * SORT_THREADS;
        if (s < valid) {
            const uint32_t kk = s_keys[s];
            const uint32_t pos = s_out_base[(kk >> shift) & mask] + s;
            keys_out[pos] = kk;
            vals_out[pos] = s_vals[s];
        }
    }
}

DigitPlan make_plan(int key_bits, int* digit_bits) {
    const int p8 = (key_bits + 7) / 8, p9 = (key_bits + 8) / 9;
    const int width = (p9 < p8) ? 9 : 8;
    DigitPlan plan;
    plan.passes = (key_bits + width - 1) / width;
    for (int p = 0; p < SORT_MAX_PASSES; p++) {
        int bits = key_bits - p * width;
        if (bits > width) bits = width;
        if (bits < 1) bits = 1;
        plan.shift[p] = p * width;
        plan.mask[p] = (1u << bits) - 1u;
    }
    *digit_bits = width;
    return plan;
}

}  // namespace

size_t sort_workspace_bytes(uint64_t n) {
    const uint64_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    // [passes][radix] histograms, [passes] tile tickets (padded), [passes][tiles][radix] status
    return sizeof(uint32_t) * (SORT_MAX_PASSES * SORT_MAX_RADIX + 64 + SORT_MAX_PASSES * tiles * SORT_MAX_RADIX);
}

int sort_num_passes(int key_bits) {
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    int width;
    return make_plan(key_bits, &width).passes;
}

// Sorts (keys_a, vals_a or iota) by the low key_bits bits.  Buffers ping-pong a -> b -> a ...;
// returns in *result_in_b whether the sorted pairs ended in the b buffers.
cudaError_t sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                       uint64_t n64, int key_bits, bool iota_vals, void* workspace,
                       int num_sms, cudaStream_t stream, bool* result_in_b) {
    *result_in_b = false;
    if (n64 == 0) return cudaSuccess;
    if (n64 >= (1ull << 30)) return cudaErrorInvalidValue;
    const uint32_t n = (uint32_t)n64;
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    int width;
    const DigitPlan plan = make_plan(key_bits, &width);
    const uint32_t radix = 1u << width;
    const uint32_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    uint32_t* hist = static_cast<uint32_t*>(workspace);          // [pass][SORT_MAX_RADIX]
    uint32_t* tickets = hist + SORT_MAX_PASSES * SORT_MAX_RADIX;
    uint32_t* status = tickets + 64;

    cudaError_t e = cudaMemsetAsync(workspace, 0,
        sizeof(uint32_t) * (SORT_MAX_PASSES * SORT_MAX_RADIX + 64 + (size_t)plan.passes * tiles * radix), stream);
    if (e != cudaSuccess) return e;

    const uint64_t chunk = (uint64_t)HIST_THREADS * HIST_ITEMS;
    uint32_t hist_blocks = (uint32_t)((n + chunk - 1) / chunk);
    const uint32_t max_blocks = (uint32_t)num_sms * 8;
    if (hist_blocks > max_blocks) hist_blocks = max_blocks;
    k_radix_hist<<<hist_blocks, HIST_THREADS, 0, stream>>>(keys_a, n, plan, hist);

    uint32_t* kin = keys_a; uint32_t* vin = vals_a; uint32_t* kout = keys_b; uint32_t* vout = vals_b;
    for (int p = 0; p < plan.passes; p++) {
        const uint32_t* h = hist + p * SORT_MAX_RADIX;
        uint32_t* ticket = tickets + p;
        uint32_t* st = status + (size_t)p * tiles * radix;
        const bool iota = (p == 0 && iota_vals);
#define LAUNCH(B, I) k_onesweep<B, I><<<tiles, SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, n, plan.shift[p], plan.mask[p], h, ticket, st)
        if (width == 9) { if (iota) LAUNCH(9, true); else LAUNCH(9, false); }
        else            { if (iota) LAUNCH(8, true); else LAUNCH(8, false); }
#undef LAUNCH
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    *result_in_b = (plan.passes & 1) != 0;
    return cudaGetLastError();
}

}  // namespace nprsph
