// onesweep.cu -- hand-written least-significant-digit radix sort of (cell key, slot) pairs.
//
// One histogram kernel reads the keys once and counts every 8-bit digit position; then one
// "onesweep" kernel per digit moves the pairs: each CTA takes a tile of 4096 pairs (dynamic
// tile ticket, so every predecessor tile is already resident), ranks its keys stably with
// warp match-any, publishes its per-digit counts to a tile-status array and resolves its
// global offsets by decoupled look-back over the predecessors' status words (count and flag
// share one 32-bit word, so no fence is needed).  Pairs are staged through shared memory so
// the global scatter is written in digit-contiguous runs.
//
// Only the low `key_bits` bits are sorted (ceil(log2(num_cells+1)) for cell keys).
// No reference counterpart: the reference searches neighbours all-pairs
// (rho_pres_comp.glsl:46, force_comp.glsl:48).
#include "sort.cuh"

namespace nprsph {

namespace {

constexpr uint32_t FLAG_AGG  = 1u << 30;   // tile count published
constexpr uint32_t FLAG_INCL = 1u << 31;   // inclusive prefix published
constexpr uint32_t FLAG_ANY  = FLAG_AGG | FLAG_INCL;
constexpr uint32_t VALUE_MASK = FLAG_AGG - 1;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---- digit histograms for all passes in one read of the keys --------------------------------
// Each thread walks a contiguous run of keys and run-length-compresses equal digits before
// touching shared memory: cell-ordered keys share their high digits over long runs, which
// would otherwise serialise the shared-memory atomics 32 ways.
constexpr int HIST_THREADS = 256;
constexpr int HIST_ITEMS = 16;

__global__ void __launch_bounds__(HIST_THREADS)
k_radix_hist(const uint32_t* __restrict__ keys, uint32_t n, int passes, int key_bits,
             uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[SORT_MAX_PASSES][RADIX];
    for (int i = threadIdx.x; i < SORT_MAX_PASSES * RADIX; i += HIST_THREADS) (&sh[0][0])[i] = 0;
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
        for (int p = 0; p < passes; p++) {
            const int shift = p * RADIX_BITS;
            const int bits = key_bits - shift;
            const uint32_t mask = bits >= RADIX_BITS ? (uint32_t)(RADIX - 1) : ((1u << bits) - 1u);
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
    for (int i = threadIdx.x; i < passes * RADIX; i += HIST_THREADS) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

// exclusive scan of one value per thread over a 256-thread CTA
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp_tot /*[8]*/) {
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
template <bool IOTA>
__global__ void __launch_bounds__(SORT_THREADS)
k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
           uint32_t mask, const uint32_t* __restrict__ hist, uint32_t* __restrict__ tile_counter,
           volatile uint32_t* __restrict__ status) {
    __shared__ uint32_t s_keys[SORT_TILE];
    __shared__ uint32_t s_vals[SORT_TILE];
    __shared__ uint32_t s_warp_hist[SORT_WARPS][RADIX];
    __shared__ uint32_t s_out_base[RADIX];    // global position of local slot 0 of each digit run
    __shared__ uint32_t s_local_off[RADIX];
    __shared__ uint32_t s_scan[SORT_WARPS];
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&s_warp_hist[0][0])[i] = 0;
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

    // stable rank of every key among the keys of its warp with the same digit
    // All match-any votes are issued first (independent), then one shared-memory atomic per
    // distinct digit per row (by the lowest peer lane) claims the rank base, which is broadcast
    // back with a shuffle.  Atomics of one warp to one address retire in program order, so row k
    // ranks below row k+1: the sort stays stable, and the 16 rows overlap instead of forming one
    // load -> store dependency chain.
    uint32_t rank[SORT_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* my_hist = s_warp_hist[warp];
    // Cell-ordered keys share their upper digits over long runs: when a whole row of 32 keys has
    // one digit (two REDUX votes), the match-any vote -- the slowest instruction of the kernel --
    // is skipped.
    uint32_t peers[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        // MATCH.ANY costs one round per distinct value in the warp.  Cell-ordered keys give rows
        // whose digits are either all equal or count upwards (consecutive cells; after the first
        // pass, keys 256 apart), i.e. all distinct: both cases are recognised with two shuffles
        // and a vote, and only irregular rows pay for the match.
        const uint32_t d = (key[k] >> shift) & mask;
        const uint32_t e = (d - __shfl_sync(0xffffffffu, d, 0)) & mask;     // offset from lane 0, cyclic
        const uint32_t below_e = __shfl_up_sync(0xffffffffu, e, 1);
        const bool rising = lane == 0 || e > below_e;
        if (!__any_sync(0xffffffffu, e != 0u))      peers[k] = 0xffffffffu;           // one digit
        else if (__all_sync(0xffffffffu, rising))   peers[k] = 1u << lane;            // all distinct
        else                                        peers[k] = __match_any_sync(0xffffffffu, d);
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

    // thread d owns digit d: counts across warps -> exclusive warp offsets + tile count
    const uint32_t d_own = tid;
    uint32_t tile_count = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
        const uint32_t c = s_warp_hist[w][d_own];
        s_warp_hist[w][d_own] = tile_count;
        tile_count += c;
    }
    status[(uint64_t)tile * RADIX + d_own] = tile_count | (tile == 0 ? FLAG_INCL : FLAG_AGG);

    const uint32_t global_excl = block_excl_scan_256(hist[d_own], s_scan);
    const uint32_t local_off = block_excl_scan_256(tile_count, s_scan);

    // decoupled look-back over predecessor tiles for this digit
    uint32_t excl = 0;
    if (tile > 0) {
        int64_t t = (int64_t)tile - 1;
        while (true) {
            const uint32_t v = status[(uint64_t)t * RADIX + d_own];
            if ((v & FLAG_ANY) == 0) continue;            // predecessor not published yet: spin
            excl += v & VALUE_MASK;
            if (v & FLAG_INCL) break;
            t--;
        }
        status[(uint64_t)tile * RADIX + d_own] = ((excl + tile_count) & VALUE_MASK) | FLAG_INCL;
    }
    s_out_base[d_own] = global_excl + excl - local_off;
    s_local_off[d_own] = local_off;
    __syncthreads();

    // stage pairs in tile-sorted order
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t d = (key[k] >> shift) & mask;
        const uint32_t p = s_local_off[d] + my_hist[d] + rank[k];
        const uint32_t idx = warp_base + k * 32 + lane;
        s_keys[p] = key[k];
        s_vals[p] = IOTA ? idx : ((idx < n) ? vals_in[idx] : 0u);
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

}  // namespace

size_t sort_workspace_bytes(uint64_t n) {
    const uint64_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    // [passes][256] histograms, [passes] tile tickets (padded), [passes][tiles][256] status
    return sizeof(uint32_t) * (SORT_MAX_PASSES * RADIX + 64 + SORT_MAX_PASSES * tiles * RADIX);
}

int sort_num_passes(int key_bits) {
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    return (key_bits + RADIX_BITS - 1) / RADIX_BITS;
}

// Sorts (keys_a, iota) by the low key_bits bits.  Buffers ping-pong a -> b -> a ...; returns in
// *result_in_b whether the sorted pairs ended in the b buffers.  vals_a is scratch.
cudaError_t sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                       uint64_t n64, int key_bits, bool iota_vals, void* workspace,
                       int num_sms, cudaStream_t stream, bool* result_in_b) {
    *result_in_b = false;
    if (n64 == 0) return cudaSuccess;
    if (n64 >= (1ull << 30)) return cudaErrorInvalidValue;
    const uint32_t n = (uint32_t)n64;
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    const int passes = sort_num_passes(key_bits);
    const uint32_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    uint32_t* hist = static_cast<uint32_t*>(workspace);
    uint32_t* tickets = hist + SORT_MAX_PASSES * RADIX;
    uint32_t* status = tickets + 64;

    cudaError_t e = cudaMemsetAsync(workspace, 0,
        sizeof(uint32_t) * (SORT_MAX_PASSES * RADIX + 64 + (size_t)passes * tiles * RADIX), stream);
    if (e != cudaSuccess) return e;

    const uint64_t chunk = (uint64_t)HIST_THREADS * HIST_ITEMS;
    uint32_t hist_blocks = (uint32_t)((n + chunk - 1) / chunk);
    const uint32_t max_blocks = (uint32_t)num_sms * 8;
    if (hist_blocks > max_blocks) hist_blocks = max_blocks;
    k_radix_hist<<<hist_blocks, HIST_THREADS, 0, stream>>>(keys_a, n, passes, key_bits, hist);

    uint32_t* kin = keys_a; uint32_t* vin = vals_a; uint32_t* kout = keys_b; uint32_t* vout = vals_b;
    for (int p = 0; p < passes; p++) {
        const int shift = p * RADIX_BITS;
        int bits = key_bits - shift; if (bits > RADIX_BITS) bits = RADIX_BITS;
        const uint32_t mask = (1u << bits) - 1u;
        if (p == 0 && iota_vals)
            k_onesweep<true><<<tiles, SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, n, shift, mask,
                hist + p * RADIX, tickets + p, status + (size_t)p * tiles * RADIX);
        else
            k_onesweep<false><<<tiles, SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, n, shift, mask,
                hist + p * RADIX, tickets + p, status + (size_t)p * tiles * RADIX);
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    *result_in_b = (passes & 1) != 0;
    return cudaGetLastError();
}

}  // namespace nprsph
