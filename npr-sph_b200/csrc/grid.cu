// grid.cu -- uniform-grid maintenance kernels and the AoS <-> cell-ordered SoA converters.
//
// HBM layout (DESIGN.md "Data layout"): the simulation state lives in cell order as three
// float4 arrays -- posid = (x, y, z, bits(original index)), velrho = (vx, vy, vz, rho),
// forcep = (fx, fy, fz, pressure) -- and is converted to/from the reference's 64-byte
// `Particle` records (Main.cpp:93-99) only at upload / publish time.
#include "kernels.cuh"

namespace nprsph {

namespace {

constexpr int TPB = 256;

inline unsigned blocks_for(uint64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

// ---- Particle[] -> SoA in original order (upload; glBufferData at Main.cpp:526) ---------------
__global__ void __launch_bounds__(TPB)
k_import(const float4* __restrict__ aos, float4* __restrict__ posid, float4* __restrict__ velrho,
         float4* __restrict__ forcep, uint32_t n) {
    const uint32_t i = blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const float4 p = aos[4 * (size_t)i + 0];
    const float4 v = aos[4 * (size_t)i + 1];
    const float4 f = aos[4 * (size_t)i + 2];
    const float4 e = aos[4 * (size_t)i + 3];
    posid[i] = make_float4(p.x, p.y, p.z, __uint_as_float(i));
    velrho[i] = make_float4(v.x, v.y, v.z, e.x);
    forcep[i] = make_float4(f.x, f.y, f.z, e.y);
}

// ---- SoA (cell order) -> Particle[] in ORIGINAL order ------------------------------------------
// Only the lanes the shaders write are stored (.xyz of pos/vel/force, extras[0..1]); the .w
// lanes and extras[2..3] keep what the caller uploaded, as in the reference.
__global__ void __launch_bounds__(TPB)
k_publish(const float4* __restrict__ posid, const float4* __restrict__ velrho,
          const float4* __restrict__ forcep, float* __restrict__ aos, uint32_t n) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    if (s >= n) return;
    const float4 p = posid[s];
    const float4 v = velrho[s];
    const float4 f = forcep[s];
    float* r = aos + 16 * (size_t)__float_as_uint(p.w);
    r[0] = p.x; r[1] = p.y; r[2] = p.z;
    r[4] = v.x; r[5] = v.y; r[6] = v.z;
    r[8] = f.x; r[9] = f.y; r[10] = f.z;
    r[12] = v.w; r[13] = f.w;
}

// ---- positions only, ORIGINAL order: what the renderer reads (attribute 0 = vec4 at offset 0 of the
// record, Main.cpp:533-535; toon_vs.glsl:17).  The w lane is the record's own (never written by a pass).
__global__ void __launch_bounds__(TPB)
k_publish_positions(const float4* __restrict__ posid, const float* __restrict__ aos,
                    float4* __restrict__ out, uint32_t n) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    if (s >= n) return;
    const float4 p = posid[s];
    const uint32_t id = __float_as_uint(p.w);
    out[id] = make_float4(p.x, p.y, p.z, aos[16 * (size_t)id + 3]);
}

// ---- host state (positions, velocities; original order) -> SoA: the inputs of a step.  Force,
// density and pressure are outputs of the passes and start from zero.
__global__ void __launch_bounds__(TPB)
k_import_state(const float4* __restrict__ pos_in, const float4* __restrict__ vel_in, uint32_t first,
               uint32_t count, float4* __restrict__ posid, float4* __restrict__ velrho,
               float4* __restrict__ forcep) {
    const uint32_t k = blockIdx.x * TPB + threadIdx.x;
    if (k >= count) return;
    const uint32_t i = first + k;
    const float4 p = pos_in[k], v = vel_in[k];
    posid[i] = make_float4(p.x, p.y, p.z, __uint_as_float(i));
    velrho[i] = make_float4(v.x, v.y, v.z, 0.0f);
    forcep[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// ---- stand-alone cell keys (only when keys are stale: after upload or a grid change) ---------
__global__ void __launch_bounds__(TPB)
k_keys(const float4* __restrict__ posid, uint32_t* __restrict__ keys, uint32_t n, GridDev g) {
    const uint32_t i = blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const float4 p = posid[i];
    keys[i] = cell_key(p.x, p.y, p.z, g);
}

// ---- gather into the new cell order + build the cell-start table -------------------------------
// cell_start[c] = first slot whose key is >= c, for c in [0, num_cells+1] (lower-bound table, so
// a run of consecutive cells [a, b] is the slot range [cell_start[a], cell_start[b+1])).
// Thread s (0..n) fills the cells between key[s-1] and key[s]; long empty runs are queued for
// k_fill_gaps so that one thread never writes millions of entries.
template <bool WITH_FORCE>
__global__ void __launch_bounds__(TPB)
k_reorder_cells(const uint32_t* __restrict__ sorted_keys, const uint32_t* __restrict__ perm,
                const float4* __restrict__ pos_in, const float4* __restrict__ vel_in,
                const float4* __restrict__ force_in, float4* __restrict__ pos_out,
                float4* __restrict__ vel_out, float4* __restrict__ force_out,
                uint32_t* __restrict__ cell_start, uint32_t num_cells, uint32_t n,
                uint4* __restrict__ gap_list, uint32_t* __restrict__ gap_count,
                uint32_t* __restrict__ tail_state) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    if (s > n) return;
    uint32_t key_here;
    if (s < n) {
        const uint32_t src = perm[s];
        key_here = sorted_keys[s];
        pos_out[s] = pos_in[src];
        vel_out[s] = vel_in[src];
        if (WITH_FORCE) force_out[s] = force_in[src];
    } else {
        key_here = num_cells + 1;                 // tail: everything above the last key -> n
    }
    const uint32_t lo = (s == 0) ? 0u : sorted_keys[s - 1] + 1u;
    if (key_here < lo) return;                    // same cell as the previous slot
    if (key_here >= num_cells && lo <= num_cells) {
        // Tail: the first slot that lies in no cell (the first NaN particle, or n when there is none)
        // owns every cell above the last real key, up to num_cells.  In a dam break that is most of
        // the table; tail_state = {F, V} records that cells F..num_cells already hold V from an
        // earlier step, so only the cells the fluid vacated since then are rewritten.
        if (s == n) cell_start[num_cells + 1u] = n;
        const uint32_t F = tail_state[0], V = tail_state[1];
        tail_state[0] = lo; tail_state[1] = s;
        key_here = num_cells;
        if (V == s) {
            if (F <= lo) return;
            if (F <= key_here) key_here = F - 1u;
        }
    }
    const uint32_t len = key_here - lo + 1u;
    if (len <= GAP_INLINE) {
        for (uint32_t c = lo; c <= key_here; c++) cell_start[c] = s;
    } else {
        push_gap(lo, len, s, gap_list, gap_count);
    }
}

// One thread block per queue entry, round robin (an evolved dam break has ~10^5 of them; with
// every block looping over the whole list this kernel took longer than the gather).
__global__ void __launch_bounds__(TPB)
k_fill_gaps(const uint4* __restrict__ gap_list, const uint32_t* __restrict__ gap_count,
            uint32_t* __restrict__ cell_start) {
    const uint32_t gaps = gap_count[0];
    for (uint32_t gi = blockIdx.x; gi < gaps; gi += gridDim.x) {
        const uint4 g = gap_list[gi];
        for (uint32_t o = threadIdx.x; o < g.y; o += TPB) cell_start[g.x + o] = g.z;
    }
}

// plain gather of the three SoA arrays by a slot permutation (snapshot restore)
__global__ void __launch_bounds__(TPB)
k_gather(const uint32_t* __restrict__ src_of_slot, const float4* __restrict__ pos_in,
         const float4* __restrict__ vel_in, const float4* __restrict__ frc_in,
         float4* __restrict__ pos_out, float4* __restrict__ vel_out, float4* __restrict__ frc_out,
         uint32_t n) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    if (s >= n) return;
    const uint32_t src = src_of_slot[s];
    pos_out[s] = pos_in[src];
    vel_out[s] = vel_in[src];
    frc_out[s] = frc_in[src];
}

__global__ void __launch_bounds__(TPB)
k_count_nan(const float4* __restrict__ posid, uint32_t n, unsigned long long* __restrict__ out) {
    const uint32_t i = blockIdx.x * TPB + threadIdx.x;
    bool bad = false;
    if (i < n) { const float4 p = posid[i]; bad = pos_is_nan(p.x, p.y, p.z); }
    const uint32_t m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

__global__ void __launch_bounds__(TPB)
k_slot_ids(const float4* __restrict__ posid, uint32_t* __restrict__ ids, uint32_t n) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    if (s < n) ids[s] = __float_as_uint(posid[s].w);
}

}  // namespace

void launch_import(const void* aos, float4* posid, float4* velrho, float4* forcep, uint32_t n,
                   cudaStream_t st) {
    if (n) k_import<<<blocks_for(n, TPB), TPB, 0, st>>>((const float4*)aos, posid, velrho, forcep, n);
}

void launch_publish(const float4* posid, const float4* velrho, const float4* forcep, void* aos,
                    uint32_t n, cudaStream_t st) {
    if (n) k_publish<<<blocks_for(n, TPB), TPB, 0, st>>>(posid, velrho, forcep, (float*)aos, n);
}

void launch_publish_positions(const float4* posid, const void* aos, float4* out, uint32_t n, cudaStream_t st) {
    if (n) k_publish_positions<<<blocks_for(n, TPB), TPB, 0, st>>>(posid, (const float*)aos, out, n);
}

void launch_import_state(const float4* pos_in, const float4* vel_in, uint32_t first, uint32_t count,
                         float4* posid, float4* velrho, float4* forcep, cudaStream_t st) {
    if (count) k_import_state<<<blocks_for(count, TPB), TPB, 0, st>>>(pos_in, vel_in, first, count, posid, velrho, forcep);
}

void launch_keys(const float4* posid, uint32_t* keys, uint32_t n, const GridDev& g, cudaStream_t st) {
    if (n) k_keys<<<blocks_for(n, TPB), TPB, 0, st>>>(posid, keys, n, g);
}

size_t gap_list_capacity(uint32_t num_cells, uint64_t n) {
    // a queued run is longer than GAP_INLINE and there is at most one run per slot (+ the tail); cutting
    // the long ones into pieces adds at most num_cells / GAP_PIECE entries
    uint64_t by_cells = ((uint64_t)num_cells + 2) / (GAP_INLINE + 1) + 2;
    uint64_t by_n = n + 1;
    return (size_t)((by_cells < by_n ? by_cells : by_n) + ((uint64_t)num_cells + 2) / GAP_PIECE + 2);
}

void launch_reorder_cells(const uint32_t* sorted_keys, const uint32_t* perm, const float4* pos_in,
                          const float4* vel_in, const float4* force_in, float4* pos_out,
                          float4* vel_out, float4* force_out, uint32_t* cell_start,
                          uint32_t num_cells, uint32_t n, uint4* gap_list, uint32_t* gap_count,
                          bool with_force, int num_sms, cudaStream_t st) {
    cudaMemsetAsync(gap_count, 0, 2 * sizeof(uint32_t), st);
    const unsigned blocks = blocks_for((uint64_t)n + 1, TPB);
    if (with_force)
        k_reorder_cells<true><<<blocks, TPB, 0, st>>>(sorted_keys, perm, pos_in, vel_in, force_in,
            pos_out, vel_out, force_out, cell_start, num_cells, n, gap_list, gap_count, gap_count + 4);
    else
        k_reorder_cells<false><<<blocks, TPB, 0, st>>>(sorted_keys, perm, pos_in, vel_in, force_in,
            pos_out, vel_out, force_out, cell_start, num_cells, n, gap_list, gap_count, gap_count + 4);
    k_fill_gaps<<<num_sms * 4, TPB, 0, st>>>(gap_list, gap_count, cell_start);
}

void launch_fill_gaps(const uint4* gap_list, const uint32_t* gap_count, uint32_t* cell_start,
                      int num_sms, cudaStream_t st) {
    k_fill_gaps<<<num_sms * 4, TPB, 0, st>>>(gap_list, gap_count, cell_start);
}

void launch_gather(const uint32_t* src_of_slot, const float4* pos_in, const float4* vel_in,
                   const float4* frc_in, float4* pos_out, float4* vel_out, float4* frc_out, uint32_t n,
                   cudaStream_t st) {
    if (n) k_gather<<<blocks_for(n, TPB), TPB, 0, st>>>(src_of_slot, pos_in, vel_in, frc_in, pos_out, vel_out, frc_out, n);
}

void launch_count_nan(const float4* posid, uint32_t n, unsigned long long* out, cudaStream_t st) {
    cudaMemsetAsync(out, 0, sizeof(unsigned long long), st);
    if (n) k_count_nan<<<blocks_for(n, TPB), TPB, 0, st>>>(posid, n, out);
}

void launch_slot_ids(const float4* posid, uint32_t* ids, uint32_t n, cudaStream_t st) {
    if (n) k_slot_ids<<<blocks_for(n, TPB), TPB, 0, st>>>(posid, ids, n);
}

}  // namespace nprsph
