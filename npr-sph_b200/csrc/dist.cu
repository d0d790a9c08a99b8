// dist.cu -- multi-GPU step: 1-D slab decomposition along x with ghost-particle halo exchange and
// particle migration (the reference itself is single-GPU, Main.cpp:668-680).
//
// One context (= one process with NCCL, or one "virtual rank" with the in-process LOCAL transport
// used by the single-GPU tests) owns the particles whose GLOBAL x cell index lies in [X0, X1).
// Slots are laid out [ghost L | own | ghost R]; because keys are x-major, the particles a
// neighbour needs as ghosts (the first / last `reach` x layers of the slab) are contiguous slot
// ranges of the sorted own array, so halo messages are sent straight out of and received
// straight into the particle arrays (no pack / unpack kernels).  Per step:
//
//   classify   own keys -> leavers copied to the migration buffers, boundary-layer counts
//   exchange   8 counters with both neighbours, ONE host sync; every later size is known
//   exchange   migrants; append immigrants (ordered by particle id) behind the own particles
//   sort       (key, slot) with the onesweep sort; leavers carry a sentinel key and drop off
//   exchange   boundary-layer positions -> ghost slots; ghost keys; cell table over all slots
//   k_rho      own slots
//   exchange   boundary-layer (velocity, rho)
//   k_force, k_integrate (+ next keys; a particle that left the slab gets GONE_L / GONE_R)
#include "dist.cuh"

#include <stdlib.h>
#include <time.h>

#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <string.h>

#include <new>

using namespace nprsph;

namespace {

// ---- NCCL through dlopen: no link-time dependency for single-GPU users --------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // a libnccl.so.2 that is already loaded (e.g. torch's bundled copy) is reused by soname
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) {
#define NCCL_SYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name)
            NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
            NCCL_SYM(CommInitRank, "ncclCommInitRank");
            NCCL_SYM(CommDestroy, "ncclCommDestroy");
            NCCL_SYM(Send, "ncclSend");
            NCCL_SYM(Recv, "ncclRecv");
            NCCL_SYM(GroupStart, "ncclGroupStart");
            NCCL_SYM(GroupEnd, "ncclGroupEnd");
            NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
            if (!api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.GroupStart ||
                !api.GroupEnd) { dlclose(api.lib); api.lib = nullptr; }
        }
    }
    return api.lib ? &api : nullptr;
}

#define NCK(ctx, call)                                                                     \
    do {                                                                                   \
        ncclResult_t r_ = (call);                                                          \
        if (r_ != ncclSuccess) {                                                           \
            (ctx)->sticky = NPRSPH_ERR_COMM;                                               \
            return fail((ctx), NPRSPH_ERR_COMM, #call ": %s",                              \
                        nccl()->GetErrorString ? nccl()->GetErrorString(r_) : "NCCL error"); \
        }                                                                                  \
    } while (0)

constexpr int TPB = 256;
inline unsigned blocks_for(uint64_t n) { return (unsigned)((n + TPB - 1) / TPB); }

// ---- kernels ---------------------------------------------------------------------------------------
// candidate particles of the global block scene for this rank: lattice planes [i0, i0+ni)
__device__ __forceinline__ uint32_t hash32(uint32_t seed, uint32_t idx) {
    uint32_t x = seed ^ (idx * 0x9E3779B9u);
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(TPB)
k_slab_scene(float4* __restrict__ posid, float4* __restrict__ velrho, uint32_t* __restrict__ keys,
             int i0, int ni, int ny, int nz, float spacing, float ox, float oy, float oz,
             float jitter, uint32_t seed, GridDev g, int W, int R) {
    const uint64_t n = (uint64_t)ni * ny * nz;
    const uint64_t t = (uint64_t)blockIdx.x * TPB + threadIdx.x;
    if (t >= n) return;
    const int k = (int)(t % nz);
    const int j = (int)((t / nz) % ny);
    const int i = i0 + (int)(t / ((uint64_t)nz * ny));
    const uint64_t idx = ((uint64_t)i * ny + j) * nz + k;          // global particle index
    float c[3] = {__fadd_rn(__fmul_rn((float)i, spacing), ox), __fadd_rn(__fmul_rn((float)j, spacing), oy),
                  __fadd_rn(__fmul_rn((float)k, spacing), oz)};
    if (jitter > 0.0f) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const uint32_t u = hash32(seed, (uint32_t)(3 * idx + a));
            const float f = __fmul_rn((float)(u >> 8), 1.0f / 16777216.0f);
            c[a] = __fadd_rn(c[a], __fmul_rn(__fsub_rn(__fmul_rn(2.0f, f), 1.0f), jitter));
        }
    }
    posid[t] = make_float4(c[0], c[1], c[2], __uint_as_float((uint32_t)idx));
    velrho[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    keys[t] = cell_key_slab(c[0], c[1], c[2], g, W, R);
}

// stand-alone classification of unsorted own keys (after a scene / upload; a step classifies
// inside k_integrate_slab)
__global__ void __launch_bounds__(TPB)
k_classify(const uint32_t* __restrict__ keys, const float4* __restrict__ posid,
           const float4* __restrict__ velrho, uint32_t n, const __grid_constant__ SlabNext sn) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    const bool live = s < n;
    const uint32_t key = live ? keys[s] : KEY_NONE;            // (out of range: none of the classes)
    const bool gone = key == KEY_GONE_L || key == KEY_GONE_R;
    const float4 p = (gone && sn.sendL) ? posid[s] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 v = (gone && sn.sendL) ? velrho[s] : make_float4(0.f, 0.f, 0.f, 0.f);
    classify_key(key, p, v, sn);
}

__global__ void __launch_bounds__(TPB)
k_migrant_ids(const Migrant* __restrict__ recv, uint32_t n, uint32_t* __restrict__ ids) {
    const uint32_t t = blockIdx.x * TPB + threadIdx.x;
    if (t < n) ids[t] = __float_as_uint(recv[t].posid.w);
}

// immigrants, in ascending particle-id order, behind the own particles; their keys
__global__ void __launch_bounds__(TPB)
k_unpack_migrants(const Migrant* __restrict__ recv, const uint32_t* __restrict__ order, uint32_t n,
                  uint32_t n_from_left, float4* __restrict__ posid, float4* __restrict__ velrho,
                  uint32_t* __restrict__ keys, GridDev g, int W, int R, uint32_t* __restrict__ errs) {
    const uint32_t t = blockIdx.x * TPB + threadIdx.x;
    if (t >= n) return;
    const uint32_t src = order ? order[t] : t;
    const Migrant m = recv[src];
    posid[t] = m.posid;
    velrho[t] = m.velrho;
    const uint32_t key = cell_key_slab(m.posid.x, m.posid.y, m.posid.z, g, W, R);
    keys[t] = key;
    // An immigrant must land in the boundary layer next to the rank it came from: the host sized
    // this step's halo messages and ghost ranges on that assumption (hL/hR in prepare_group).  A
    // particle that crossed more than `reach` cell layers in one step (|v_x| dt > reach * cell)
    // breaks it; the flag is sticky and fails the next collective call.
    const uint32_t cxl = key < KEY_NAN ? key / ((uint32_t)g.dim[1] * (uint32_t)g.dim[2]) : 0xFFFFFFFFu;
    const bool ok = src < n_from_left ? (cxl >= (uint32_t)R && cxl < (uint32_t)(2 * R))
                                      : (cxl >= (uint32_t)W && cxl < (uint32_t)(W + R));
    if (!ok) errs[ERR_IMMIGRANT] = 1u;
}

// ---- cell table of a slab ---------------------------------------------------------------------------
// Lower-bound table over the slots [ghost L | own (valid) | own (NaN) | ghost R].  Ghost-L keys lie
// in x layers [0, R), own keys in [R, R+W), ghost-R keys in [R+W, W+2R): three independent regions
// of the table.  The own region is written by the gather kernel (before any ghost has arrived),
// the two ghost regions by k_ghost_cells once the neighbours' boundary-layer positions are in.
// An element q of a sorted run [0, len] (q == len: the tail) fills the cells (key[q-1], key[q]]
// with its slot; long empty runs are queued for k_fill_gaps (grid.cu).
__device__ __forceinline__ void fill_cells(uint32_t lo, uint32_t hi, uint32_t slot,
                                           uint32_t* __restrict__ table, uint4* __restrict__ gap_list,
                                           uint32_t* __restrict__ gap_count) {
    if (hi < lo || hi == 0xFFFFFFFFu) return;
    const uint32_t len = hi - lo + 1u;
    if (len <= GAP_INLINE) {
        for (uint32_t c = lo; c <= hi; c++) table[c] = slot;
    } else {
        push_gap(lo, len, slot, gap_list, gap_count);
    }
}

// Gather of the own particles into sorted order + the own region of the cell table.
// n: own particles after migration; n_valid of them (the first) have cells, the rest are NaN.
// Cells [cell_lo, cell_hi] belong to the region; cells above the last own key hold tail_slot (the
// first ghost-R slot: the NaN block in between is harmless, a NaN candidate is never a hit).
// frc_in != nullptr (the prepare after an upload / scene): the force array follows the permutation
// too, into frc_out; sources >= n_frc (immigrants, which travel without their force) get zero.
__global__ void __launch_bounds__(TPB)
k_gather_cells_slab(const uint32_t* __restrict__ sorted_keys, const uint32_t* __restrict__ perm,
                    const float4* __restrict__ pos_in, const float4* __restrict__ vel_in,
                    float4* __restrict__ pos_out, float4* __restrict__ vel_out,
                    const float4* __restrict__ frc_in, float4* __restrict__ frc_out, uint32_t n_frc, uint32_t n,
                    uint32_t n_valid, uint32_t slot_base, uint32_t tail_slot, uint32_t cell_lo,
                    uint32_t cell_hi, uint32_t* __restrict__ table, uint4* __restrict__ gap_list,
                    uint32_t* __restrict__ gap_count, uint32_t plane, uint32_t* __restrict__ xmax_out) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    // top occupied x layer of this arrangement (keys are x-major): bounds next step's sort and table
    if (s == 0) *xmax_out = n_valid ? sorted_keys[n_valid - 1u] / plane : 0u;
    if (s < n) {
        const uint32_t src = perm[s];
        pos_out[s] = pos_in[src];
        vel_out[s] = vel_in[src];
        if (frc_in) frc_out[s] = src < n_frc ? frc_in[src] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (s > n_valid) return;
    const uint32_t lo = (s == 0) ? cell_lo : max(sorted_keys[s - 1] + 1u, cell_lo);
    const uint32_t hi = (s < n_valid) ? min(sorted_keys[s], cell_hi) : cell_hi;
    fill_cells(lo, hi, (s < n_valid) ? slot_base + s : tail_slot, table, gap_list, gap_count);
}

// One ghost block: keys from the received positions (the same arithmetic every rank uses for a
// cell) + its region [cell_lo, cell_hi] of the table; cells above the last ghost key hold tail_slot.
__global__ void __launch_bounds__(TPB)
k_ghost_cells(const float4* __restrict__ posid, uint32_t n, uint32_t slot_base, uint32_t tail_slot,
              uint32_t cell_lo, uint32_t cell_hi, GridDev g, uint32_t* __restrict__ table,
              uint4* __restrict__ gap_list, uint32_t* __restrict__ gap_count) {
    const uint32_t q = blockIdx.x * TPB + threadIdx.x;
    if (q > n) return;
    uint32_t prev = 0u, here = 0u;
    if (q > 0) { const float4 p = posid[q - 1]; prev = cell_key(p.x, p.y, p.z, g); }
    if (q < n) { const float4 p = posid[q]; here = cell_key(p.x, p.y, p.z, g); }
    const uint32_t lo = (q == 0) ? cell_lo : max(prev + 1u, cell_lo);
    const uint32_t hi = (q < n) ? min(here, cell_hi) : cell_hi;
    fill_cells(lo, hi, (q < n) ? slot_base + q : tail_slot, table, gap_list, gap_count);
}

__global__ void __launch_bounds__(TPB)
k_integrate_slab(float4* __restrict__ posid, float4* __restrict__ velrho,
                 const float4* __restrict__ forcep, uint32_t* __restrict__ keys, uint32_t n,
                 SphDev sp, const __grid_constant__ ColliderSet cs, const __grid_constant__ SlabNext sn) {
    const uint32_t i = blockIdx.x * TPB + threadIdx.x;
    uint32_t key = KEY_NONE;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), v = p;
    if (i < n) {
        p = posid[i];
        v = velrho[i];
        const float4 f = forcep[i];
        integrate_particle(p, v, f, sp, cs);
        posid[i] = p;
        velrho[i] = v;
        key = cell_key_slab(p.x, p.y, p.z, sn.g, sn.W, sn.R);
        keys[i] = key;
    }
    // the next step's classification (leavers, boundary layers, NaN, top layer) while the particle
    // is in registers
    classify_key(key, p, v, sn);
}

__global__ void k_set_words(uint32_t* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) {
    dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d; dst[4] = e;
}

__global__ void __launch_bounds__(TPB)
k_pack_records(const float4* __restrict__ posid, const float4* __restrict__ velrho,
               const float4* __restrict__ forcep, uint32_t n, float4* __restrict__ rec,
               uint32_t* __restrict__ ids) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    if (s >= n) return;
    const float4 p = posid[s], v = velrho[s], f = forcep[s];
    rec[4 * (size_t)s + 0] = make_float4(p.x, p.y, p.z, 1.0f);
    rec[4 * (size_t)s + 1] = make_float4(v.x, v.y, v.z, 0.0f);
    rec[4 * (size_t)s + 2] = make_float4(f.x, f.y, f.z, 0.0f);
    rec[4 * (size_t)s + 3] = make_float4(v.w, f.w, 0.0f, 0.0f);
    ids[s] = __float_as_uint(p.w);
}

// host records + ids (nprsph_dist_upload) -> own slots and slab keys
__global__ void __launch_bounds__(TPB)
k_unpack_records(const float4* __restrict__ rec, const uint32_t* __restrict__ ids, uint32_t n,
                 float4* __restrict__ posid, float4* __restrict__ velrho, float4* __restrict__ forcep,
                 uint32_t* __restrict__ keys, GridDev g, int W, int R) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    if (s >= n) return;
    const float4 p = rec[4 * (size_t)s], v = rec[4 * (size_t)s + 1], f = rec[4 * (size_t)s + 2];
    const float4 e = rec[4 * (size_t)s + 3];
    posid[s] = make_float4(p.x, p.y, p.z, __uint_as_float(ids[s]));
    velrho[s] = make_float4(v.x, v.y, v.z, e.x);
    forcep[s] = make_float4(f.x, f.y, f.z, e.y);
    keys[s] = cell_key_slab(p.x, p.y, p.z, g, W, R);
}

// staged inputs of a step (nprsph_dist_upload_state) -> own slots + slab keys
__global__ void __launch_bounds__(TPB)
k_import_state_slab(const float4* __restrict__ spos, const float4* __restrict__ svel, uint32_t n,
                    float4* __restrict__ posid, float4* __restrict__ velrho, uint32_t* __restrict__ keys,
                    GridDev g, int W, int R) {
    const uint32_t s = blockIdx.x * TPB + threadIdx.x;
    if (s >= n) return;
    const float4 p = spos[s];
    posid[s] = p;
    velrho[s] = svel[s];
    keys[s] = cell_key_slab(p.x, p.y, p.z, g, W, R);
}

// ---- neighbour exchange --------------------------------------------------------------------------
// Sends sendL/sendR to the left/right rank and receives recvL/recvR from them (byte counts; a
// rank at the end of the chain has no neighbour on that side).  LOCAL transport: the "ranks" are
// contexts of this process sharing one stream, so a receive is a device copy out of the
// neighbour's send pointer; the group driver posts all sends before any receive runs.
struct Posted { const void* toL = nullptr; const void* toR = nullptr; };

int exchange_nccl(nprsph_ctx* c, cudaStream_t st, const void* sendL, size_t nL, const void* sendR,
                  size_t nR, void* recvL, size_t rL, void* recvR, size_t rR) {
    DistState* d = c->dist;
    NcclApi* api = nccl();
    ncclComm_t comm = (ncclComm_t)d->nccl_comm;
    NCK(c, api->GroupStart());
    if (d->rank > 0) {
        if (nL) NCK(c, api->Send(sendL, nL, ncclChar, d->rank - 1, comm, st));
        if (rL) NCK(c, api->Recv(recvL, rL, ncclChar, d->rank - 1, comm, st));
    }
    if (d->rank + 1 < d->world) {
        if (nR) NCK(c, api->Send(sendR, nR, ncclChar, d->rank + 1, comm, st));
        if (rR) NCK(c, api->Recv(recvR, rR, ncclChar, d->rank + 1, comm, st));
    }
    NCK(c, api->GroupEnd());
    return NPRSPH_OK;
}

struct Xfer { const void* sendL; size_t nL; const void* sendR; size_t nR; void* recvL; size_t rL; void* recvR; size_t rR; };

// on_comm: run the transfers on each rank's communication stream (the caller orders it against
// the compute stream with events) instead of the compute stream
int exchange_group(nprsph_ctx** cs, int n, const Xfer* x, bool on_comm = false) {
    auto stream_of = [&](nprsph_ctx* c) { return on_comm ? c->dist->comm_stream : c->stream; };
    if (n == 1 && cs[0]->dist->transport == NPRSPH_TRANSPORT_NCCL)
        return exchange_nccl(cs[0], stream_of(cs[0]), x[0].sendL, x[0].nL, x[0].sendR, x[0].nR, x[0].recvL,
                             x[0].rL, x[0].recvR, x[0].rR);
    for (int r = 0; r < n; r++) {          // LOCAL: rank r receives from r-1 (its sendR) and r+1 (its sendL)
        nprsph_ctx* c = cs[r];
        if (r > 0 && x[r].rL) {
            if (x[r].rL != x[r - 1].nR) return fail(c, NPRSPH_ERR_COMM, "local exchange size mismatch%s");
            CK(c, cudaMemcpyAsync(x[r].recvL, x[r - 1].sendR, x[r].rL, cudaMemcpyDeviceToDevice, stream_of(c)));
        }
        if (r + 1 < n && x[r].rR) {
            if (x[r].rR != x[r + 1].nL) return fail(c, NPRSPH_ERR_COMM, "local exchange size mismatch%s");
            CK(c, cudaMemcpyAsync(x[r].recvR, x[r + 1].sendL, x[r].rR, cudaMemcpyDeviceToDevice, stream_of(c)));
        }
    }
    return NPRSPH_OK;
}

// ---- host-side phases -------------------------------------------------------------------------------
// local grid of the slab [X0, X0 + W): own layers plus R ghost layers on each side, addressed with
// global cell arithmetic minus the integer offset x_off
GridDev local_grid_of(const nprsph_ctx* c, int X0, int W, int R) {
    GridDev g = c->grid;
    g.x_off = X0 - R;
    g.dimx_global = c->grid.dim[0];
    g.dim[0] = W + 2 * R;
    g.num_cells = (uint32_t)((uint64_t)g.dim[0] * g.dim[1] * g.dim[2]);
    return g;
}

int setup_local_grid(nprsph_ctx* c) {
    DistState* d = c->dist;
    int rc = refresh_params(c);
    if (rc) return rc;
    d->R = c->grid.reach;
    d->W = d->X1 - d->X0;
    const uint64_t cells = (uint64_t)(d->W + 2 * d->R) * c->grid.dim[1] * c->grid.dim[2];
    if (cells + 4 >= (1ull << 32)) return fail(c, NPRSPH_ERR_INVALID, "local grid too large%s");
    d->lg = local_grid_of(c, d->X0, d->W, d->R);
    int bits = 1;
    while (bits < 32 && (1ull << bits) < cells + 4) bits++;      // cell keys below 2^bits - 4 (sentinels)
    d->key_bits = bits;
    const size_t need = (size_t)cells + 4;
    if (need > c->cell_cap) {
        // (re-balancing widens a slab one layer at a time: leave room for a quarter more layers)
        const size_t slack = d->rebalance_every ? (size_t)(d->W / 4 + 2) * c->grid.dim[1] * c->grid.dim[2] : 0;
        size_t want = need + slack;
        if (want + 4 >= ((size_t)1 << 32)) want = need;
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, realloc_dev(c->cell_start, want));
        c->cell_cap = want;
        CK(c, cudaMemsetAsync(c->cell_start, 0, want * sizeof(uint32_t), c->stream));
    }
    const size_t gaps = gap_list_capacity((uint32_t)cells + 2u, d->cap_total ? d->cap_total : 1);
    if (gaps > c->gap_cap) {
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, realloc_dev(c->gap_list, gaps + gaps / 4));
        c->gap_cap = gaps + gaps / 4;
    }
    return NPRSPH_OK;
}

// what the key-producing kernels of this step need (slab.cuh): the faces the NEXT prepare will use
SlabNext slab_next(const nprsph_ctx* c, bool with_send) {
    const DistState* d = c->dist;
    SlabNext sn;
    sn.W = d->X1_next - d->X0_next;
    sn.R = d->R;
    sn.g = local_grid_of(c, d->X0_next, sn.W, d->R);
    sn.sendL = with_send ? d->sendL : nullptr;
    sn.sendR = with_send ? d->sendR : nullptr;
    sn.cap_mig = d->cap_mig;
    sn.counts = d->d_counts;
    sn.errs = d->d_counts + 3 * CNT_WORDS;
    return sn;
}

// zero this step's counters and publish the host-known ones the neighbours need for re-balancing
int reset_counts(nprsph_ctx* c) {
    DistState* d = c->dist;
    CK(c, cudaMemsetAsync(d->d_counts, 0, 3 * CNT_WORDS * sizeof(uint32_t), c->stream));
    k_set_words<<<1, 1, 0, c->stream>>>(d->d_counts + CNT_NOWN, d->n_own, d->cap_own - d->n_own,
                                         (uint32_t)(d->X1_next - d->X0_next), d->cap_mig,
                                         (uint32_t)(d->cost_ms * 1000.0f));
    return NPRSPH_OK;
}

int alloc_slab(nprsph_ctx* c, uint64_t cap_own, uint64_t cap_ghost, uint64_t cap_mig) {
    DistState* d = c->dist;
    cap_ghost = (cap_ghost + 1) & ~(uint64_t)1;      // own_off even: slot pairs are (even, odd)
    const uint64_t total = cap_own + 2 * cap_ghost;
    if (total >= (1ull << 30)) return fail(c, NPRSPH_ERR_INVALID, "slab capacity too large%s");
    CK(c, cudaStreamSynchronize(c->stream));
    { float4* a = (float4*)c->aos; CK(c, realloc_dev(a, total * 4)); c->aos = a; }
    for (int b = 0; b < 2; b++) {
        CK(c, realloc_dev(c->pos[b], total));
        CK(c, realloc_dev(c->vel[b], total));
        CK(c, realloc_dev(c->keys[b], total));
        CK(c, realloc_dev(c->vals[b], total));
    }
    CK(c, realloc_dev(c->frc[0], total));
    CK(c, realloc_dev(c->frc[1], (size_t)0));
    { char* w = (char*)c->sort_ws; CK(c, realloc_dev(w, sort_workspace_bytes(total))); c->sort_ws = w; }
    { int rc = ensure_records(c, total, c->grid.reach); if (rc) return rc; }
    if (c->cfg.flags & NPRSPH_FLAG_COUNT_NEIGHBOURS) {
        CK(c, realloc_dev(c->counts_rho, total));
        CK(c, realloc_dev(c->counts_force, total));
    }
    CK(c, realloc_dev(d->sendL, cap_mig));
    CK(c, realloc_dev(d->sendR, cap_mig));
    CK(c, realloc_dev(d->recv, 2 * cap_mig));
    CK(c, realloc_dev(d->mig_ids, 8 * cap_mig));
    { char* w = (char*)d->mig_sort_ws; CK(c, realloc_dev(w, sort_workspace_bytes(2 * cap_mig))); d->mig_sort_ws = w; }
    c->cap = total;
    c->n = 0;
    d->cap_own = (uint32_t)cap_own; d->cap_ghost = (uint32_t)cap_ghost; d->cap_mig = (uint32_t)cap_mig;
    d->cap_total = (uint32_t)total; d->own_off = (uint32_t)cap_ghost;
    return NPRSPH_OK;
}

// NPRSPH_DIST_TRACE=1: synchronise after every phase of prepare_group and report the mean wall time
// of each on stderr every 20 steps (a diagnostic: the synchronisation serialises the step).
struct PhaseTrace {
    bool on = getenv("NPRSPH_DIST_TRACE") != nullptr;
    double acc[8] = {0}, t0 = 0;
    int steps = 0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    void begin(nprsph_ctx** cs, int n) { if (!on) return; for (int r = 0; r < n; r++) cudaStreamSynchronize(cs[r]->stream); t0 = now(); }
    void mark(nprsph_ctx** cs, int n, int phase) {
        if (!on) return;
        for (int r = 0; r < n; r++) cudaStreamSynchronize(cs[r]->stream);
        const double t = now(); acc[phase] += t - t0; t0 = t;
    }
    void end(int rank) {
        if (!on || ++steps % 20) return;
        fprintf(stderr, "[nprsph dist trace] rank %d, mean ms over %d steps: classify+counts+sync %.3f | migrants %.3f | "
                "unpack+sort %.3f | gather+own cells %.3f | ghost positions %.3f | ghost cells+gaps %.3f\n", rank, steps,
                acc[0] / steps, acc[1] / steps, acc[2] / steps, acc[3] / steps, acc[4] / steps, acc[5] / steps);
    }
};
static PhaseTrace g_trace;

// Re-balancing: every `rebalance_every` steps each interior slab face may move by
// ONE x layer towards the lighter rank.  Both ranks of a face evaluate this function on the same
// numbers (their two counter blocks, exchanged in phase (1)), so they take the same decision
// without talking to anyone else; the layer that changes hands simply becomes "gone" under the
// next local grid and travels through the ordinary migration path of the next prepare.
// a = counters of the left rank of the face, b = of the right rank.  Returns -1 (face moves left:
// a gives its last layer to b), +1 (b gives its first layer to a) or 0.
// by_time: the weight of a rank is the measured duration of its density pass (CNT_COST, microseconds)
// instead of its particle count -- the front of a dam break costs more per particle than the bulk.
int face_move(const uint32_t* a, const uint32_t* b, int R, uint32_t cap_ghost, bool by_time) {
    // With reach 1 (cell = h) the receiver's boundary layer is ONE cell layer: the layer handed over
    // fills it, and a particle that crosses the old face in the same step would land one layer deeper
    // than the halo accounting allows.  Faces only move when the boundary layer has room for both.
    if (R < 2) return 0;
    const int64_t nA = a[CNT_NOWN], nB = b[CNT_NOWN];
    const int64_t layerA = (a[CNT_HALO_R] + R - 1) / R, layerB = (b[CNT_HALO_L] + R - 1) / R;   // particles per x layer at the face
    if (by_time && a[CNT_COST] && b[CNT_COST] && nA > 0 && nB > 0) {
        const int64_t wA = a[CNT_COST], wB = b[CNT_COST];
        const int64_t lwA = wA * layerA / nA, lwB = wB * layerB / nB;       // cost of the layer at the face
        const int64_t cap_mig = a[CNT_CAPMIG] < b[CNT_CAPMIG] ? a[CNT_CAPMIG] : b[CNT_CAPMIG];
        const int wmin = 2 * R + 2;
        auto fits = [&](int64_t layer, int64_t free_rx) {
            return 2 * layer + 1024 <= cap_mig && 2 * layer + 65536 <= free_rx &&
                   (int64_t)(R + 2) * layer * 3 / 2 <= (int64_t)cap_ghost;
        };
        // (3 layers of hysteresis: the times are measurements)
        if (wA > wB + 3 * lwA + 1 && (int)a[CNT_WIDTH] >= wmin && fits(layerA, b[CNT_FREE])) return -1;
        if (wB > wA + 3 * lwB + 1 && (int)b[CNT_WIDTH] >= wmin && fits(layerB, a[CNT_FREE])) return +1;
        return 0;
    }
    const int64_t cap_mig = a[CNT_CAPMIG] < b[CNT_CAPMIG] ? a[CNT_CAPMIG] : b[CNT_CAPMIG];
    const int wmin = 2 * R + 2;           // a slab keeps 2R layers after giving one away on EACH face
    auto fits = [&](int64_t layer, int64_t free_rx) {
        // the hand-over rides on top of the ordinary migrants and becomes ghosts of the giver
        return 2 * layer + 1024 <= cap_mig && 2 * layer + 65536 <= free_rx &&
               (int64_t)(R + 2) * layer * 3 / 2 <= (int64_t)cap_ghost;
    };
    if (nA > nB + 2 * layerA && (int)a[CNT_WIDTH] >= wmin && fits(layerA, b[CNT_FREE])) return -1;
    if (nB > nA + 2 * layerB && (int)b[CNT_WIDTH] >= wmin && fits(layerB, a[CNT_FREE])) return +1;
    return 0;
}

// Everything between "positions + unsorted slab keys of the own particles are final" and "cell
// table and ghosts are ready for k_rho", for all local ranks in lock step.
int prepare_group(nprsph_ctx** cs, int n) {
    Xfer x[64] = {};
    g_trace.begin(cs, n);
    // (0) the faces decided one step ago take effect: the keys at hand were computed against them
    for (int r = 0; r < n; r++) {
        nprsph_ctx* c = cs[r]; DistState* d = c->dist;
        CK(c, cudaSetDevice(c->cfg.device));
        if (d->X0_next != d->X0 || d->X1_next != d->X1) {
            d->X0 = d->X0_next; d->X1 = d->X1_next;
            int rc = setup_local_grid(c);
            if (rc) return rc;
        }
    }
    // (1) classify + counter exchange, one host sync
    for (int r = 0; r < n; r++) {
        nprsph_ctx* c = cs[r]; DistState* d = c->dist;
        CK(c, cudaSetDevice(c->cfg.device));
        if (!d->classified) {
            int rc = reset_counts(c);
            if (rc) return rc;
            if (d->n_own)
                k_classify<<<blocks_for(d->n_own), TPB, 0, c->stream>>>(
                    c->keys[0], c->pos[c->cur] + d->own_off, c->vel[c->cur] + d->own_off, d->n_own,
                    slab_next(c, !d->first_prepare));
        }
        d->classified = false;
        x[r] = {d->d_counts, CNT_WORDS * 4, d->d_counts, CNT_WORDS * 4,
                d->d_counts + CNT_WORDS, CNT_WORDS * 4, d->d_counts + 2 * CNT_WORDS, CNT_WORDS * 4};
    }
    int rc = exchange_group(cs, n, x);
    if (rc) return rc;
    for (int r = 0; r < n; r++) {
        nprsph_ctx* c = cs[r];
        CK(c, cudaMemcpyAsync(c->dist->h_counts, c->dist->d_counts, (3 * CNT_WORDS + ERR_WORDS) * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, c->stream));
    }
    for (int r = 0; r < n; r++) CK(cs[r], cudaStreamSynchronize(cs[r]->stream));
    for (int r = 0; r < n; r++) {          // duration of the last density pass (its events have completed)
        DistState* d = cs[r]->dist;
        float ms = 0.f;
        if (d->work_timed && cudaEventElapsedTime(&ms, d->ev_work0, d->ev_work1) == cudaSuccess)
            d->cost_ms = d->cost_ms > 0.f ? 0.75f * d->cost_ms + 0.25f * ms : ms;
    }
    g_trace.mark(cs, n, 0);

    // (2) sizes, migrant exchange
    uint32_t inL[64], inR[64], leaveL[64], leaveR[64], dropped[64];
    for (int r = 0; r < n; r++) {
        nprsph_ctx* c = cs[r]; DistState* d = c->dist;
        const uint32_t* mine = d->h_counts;
        const uint32_t* fromL = d->h_counts + CNT_WORDS;
        const uint32_t* fromR = d->h_counts + 2 * CNT_WORDS;
        const uint32_t* errs = d->h_counts + 3 * CNT_WORDS;
        const bool hasL = d->rank > 0, hasR = d->rank + 1 < d->world;
        if (errs[ERR_OVERFLOW]) return fail(c, NPRSPH_ERR_NOMEM, "migration buffer overflow (raise max_migrate)%s");
        if (errs[ERR_IMMIGRANT])
            return fail(c, NPRSPH_ERR_STATE, "a particle crossed more than `reach` cell layers of a slab face in one step "
                        "(|v_x| dt > reach * cell): halo sizes of the last step were wrong%s");
        dropped[r] = mine[CNT_LEAVE_L] + mine[CNT_LEAVE_R];
        const bool fp = d->first_prepare;
        leaveL[r] = fp ? 0u : mine[CNT_LEAVE_L];
        leaveR[r] = fp ? 0u : mine[CNT_LEAVE_R];
        if (!fp && ((!hasL && leaveL[r]) || (!hasR && leaveR[r])))
            return fail(c, NPRSPH_ERR_STATE, "particle left the global grid through a slab face%s");
        inL[r] = (hasL && !fp) ? fromL[CNT_LEAVE_R] : 0u;
        inR[r] = (hasR && !fp) ? fromR[CNT_LEAVE_L] : 0u;
        if (inL[r] + inR[r] > 2 * d->cap_mig || d->n_own + inL[r] + inR[r] > d->cap_own)
            return fail(c, NPRSPH_ERR_NOMEM, "own-particle capacity exceeded (raise max_own)%s");
        d->n_nan = mine[CNT_NAN];
        {   // x layers the table and the sort must cover.  A rank with a right neighbour: all of them
            // (ghosts and immigrants sit at the top).  The last rank owns the long empty stretch of box
            // ahead of the dam break: the top occupied layer of the LAST sort (read off its last key by
            // the gather kernel) + `reach` layers a particle may have moved since + `reach` layers the
            // walks look ahead + 1 for a face that moved in between.
            uint32_t top = (uint32_t)d->lg.dim[0] - 1u;
            if (!hasR && d->xmax_known && !fp) {
                const uint32_t t = errs[STICKY_XMAX] + 2u * (uint32_t)d->R + 1u;
                if (t < top) top = t;
            }
            d->x_top = top;
        }
        // boundary layers after migration: stayers + the immigrants that arrive through that face
        d->hL = mine[CNT_HALO_L] + inL[r];
        d->hR = mine[CNT_HALO_R] + inR[r];
        d->gL = hasL ? fromL[CNT_HALO_R] + leaveL[r] : 0u;
        d->gR = hasR ? fromR[CNT_HALO_L] + leaveR[r] : 0u;
        if (d->gL > d->cap_ghost || d->gR > d->cap_ghost)
            return fail(c, NPRSPH_ERR_NOMEM, "ghost capacity exceeded (raise max_ghost)%s");
        d->migrated_total += leaveL[r] + leaveR[r];
        d->last_migrated = leaveL[r] + leaveR[r];
        // faces of the prepare after this one (the coming step keys its particles against them)
        if (d->rebalance_every > 0 && !d->faces_frozen && !fp && d->steps_done > 0 &&
            d->steps_done % (uint64_t)d->rebalance_every == 0) {
            const int mL = hasL ? face_move(fromL, mine, d->R, d->cap_ghost, d->balance_time) : 0;
            const int mR = hasR ? face_move(mine, fromR, d->R, d->cap_ghost, d->balance_time) : 0;
            d->X0_next = d->X0 + mL;
            d->X1_next = d->X1 + mR;
            d->rebalanced += (uint64_t)((mL != 0) + (mR != 0));
        }
        x[r] = {d->sendL, leaveL[r] * sizeof(Migrant), d->sendR, leaveR[r] * sizeof(Migrant),
                d->recv, inL[r] * sizeof(Migrant), d->recv + inL[r], inR[r] * sizeof(Migrant)};
    }
    rc = exchange_group(cs, n, x);
    if (rc) return rc;
    g_trace.mark(cs, n, 1);

    // (3) immigrants behind the own particles (ascending id), sort, gather into [own_off, ...)
    for (int r = 0; r < n; r++) {
        nprsph_ctx* c = cs[r]; DistState* d = c->dist;
        CK(c, cudaSetDevice(c->cfg.device));
        const uint32_t n_in = inL[r] + inR[r];
        const uint32_t n_pre = d->n_own + n_in;
        float4* pos = c->pos[c->cur] + d->own_off;
        float4* vel = c->vel[c->cur] + d->own_off;
        if (n_in) {
            uint32_t* ids = d->mig_ids;
            uint32_t* order = nullptr;
            if (n_in > 1) {
                k_migrant_ids<<<blocks_for(n_in), TPB, 0, c->stream>>>(d->recv, n_in, ids);
                bool in_b = false;
                CK(c, sort_pairs(ids, ids + 2 * d->cap_mig, ids + 4 * d->cap_mig, ids + 6 * d->cap_mig, n_in,
                                 32, true, d->mig_sort_ws, c->num_sms, c->stream, &in_b));
                order = in_b ? ids + 6 * d->cap_mig : ids + 2 * d->cap_mig;
            }
            k_unpack_migrants<<<blocks_for(n_in), TPB, 0, c->stream>>>(
                d->recv, order, n_in, inL[r], pos + d->n_own, vel + d->n_own, c->keys[0] + d->n_own, d->lg, d->W,
                d->R, d->d_counts + 3 * CNT_WORDS);
        }
        const uint32_t n_new = n_pre - (d->first_prepare ? dropped[r] : leaveL[r] + leaveR[r]);
        bool in_b = false;
        // digit passes over the occupied x layers only (KEY_* sentinels sort behind them)
        int bits = 1;
        const uint64_t span = (uint64_t)(d->x_top + 1u) * (uint32_t)d->lg.dim[1] * (uint32_t)d->lg.dim[2] + 4u;
        while (bits < d->key_bits && (1ull << bits) < span) bits++;
        d->sort_bits = bits;
        CK(c, sort_pairs(c->keys[0], c->vals[0], c->keys[1], c->vals[1], n_pre, bits, true,
                         c->sort_ws, c->num_sms, c->stream, &in_b));
        if (g_trace.on && r == n - 1) g_trace.mark(cs, n, 2);
        const int nxt = 1 - c->cur;
        {   // gather + own region of the cell table (layers [R, R+W), up to the last layer in use)
            const uint32_t plane = (uint32_t)d->lg.dim[1] * (uint32_t)d->lg.dim[2];
            const uint32_t table_len = (d->x_top + 1u) * plane + 4u;
            const uint32_t own_hi = ((d->rank + 1 < d->world) ? (uint32_t)(d->R + d->W) * plane : table_len) - 1u;
            CK(c, cudaMemsetAsync(c->gap_count, 0, 2 * sizeof(uint32_t), c->stream));
            // after an upload / scene the force array follows the permutation too (through the idle
            // AoS staging buffer): later steps recompute it before anyone can read it
            float4* frc_tmp = d->gather_force ? (float4*)c->aos : nullptr;
            k_gather_cells_slab<<<blocks_for((uint64_t)n_new + 1), TPB, 0, c->stream>>>(
                in_b ? c->keys[1] : c->keys[0], in_b ? c->vals[1] : c->vals[0], pos, vel,
                c->pos[nxt] + d->own_off, c->vel[nxt] + d->own_off,
                frc_tmp ? c->frc[0] + d->own_off : nullptr, frc_tmp, d->n_own, n_new,
                n_new - d->n_nan, d->own_off,
                d->own_off + n_new, (uint32_t)d->R * plane, own_hi, c->cell_start, c->gap_list, c->gap_count,
                plane, d->d_counts + 3 * CNT_WORDS + STICKY_XMAX);
            d->xmax_known = true;
            if (frc_tmp && n_new)
                CK(c, cudaMemcpyAsync(c->frc[0] + d->own_off, frc_tmp, (size_t)n_new * sizeof(float4),
                                      cudaMemcpyDeviceToDevice, c->stream));
            d->gather_force = false;
        }
        c->cur = nxt;
        d->n_own = n_new;
        d->first_prepare = false;
        // (4) boundary-layer positions -> neighbours' ghost slots
        float4* p = c->pos[c->cur];
        const uint32_t right0 = d->own_off + d->n_own - d->n_nan - d->hR;
        x[r] = {p + d->own_off, (d->rank > 0 ? d->hL : 0u) * sizeof(float4),
                p + right0, (d->rank + 1 < d->world ? d->hR : 0u) * sizeof(float4),
                p + d->own_off - d->gL, d->gL * sizeof(float4),
                p + d->own_off + d->n_own, d->gR * sizeof(float4)};
    }
    g_trace.mark(cs, n, 3);
    rc = exchange_group(cs, n, x);
    if (rc) return rc;
    g_trace.mark(cs, n, 4);

    // (5) ghost regions of the cell table (layers [0, R) and [R+W, W+2R)), then the long empty runs
    for (int r = 0; r < n; r++) {
        nprsph_ctx* c = cs[r]; DistState* d = c->dist;
        CK(c, cudaSetDevice(c->cfg.device));
        const float4* p = c->pos[c->cur];
        const uint32_t plane = (uint32_t)d->lg.dim[1] * (uint32_t)d->lg.dim[2];
        const uint32_t table_len = (d->x_top + 1u) * plane + 4u;
        // left: always written (without a left neighbour the layers [0, R) are empty and hold own_off)
        k_ghost_cells<<<blocks_for((uint64_t)d->gL + 1), TPB, 0, c->stream>>>(
            p + d->own_off - d->gL, d->gL, d->own_off - d->gL, d->own_off, 0u, (uint32_t)d->R * plane - 1u,
            d->lg, c->cell_start, c->gap_list, c->gap_count);
        if (d->rank + 1 < d->world)
            k_ghost_cells<<<blocks_for((uint64_t)d->gR + 1), TPB, 0, c->stream>>>(
                p + d->own_off + d->n_own, d->gR, d->own_off + d->n_own, d->own_off + d->n_own + d->gR,
                (uint32_t)(d->R + d->W) * plane, table_len - 1u, d->lg, c->cell_start, c->gap_list, c->gap_count);
        launch_fill_gaps(c->gap_list, c->gap_count, c->cell_start, c->num_sms, c->stream);
        CK(c, cudaGetLastError());
    }
    g_trace.mark(cs, n, 5);
    g_trace.end(cs[0]->dist->rank);
    return NPRSPH_OK;
}

int ensure_prepared(nprsph_ctx** cs, int n) {
    bool all = true;
    for (int r = 0; r < n; r++) all = all && cs[r]->dist->prepared;
    if (all) return NPRSPH_OK;
    int rc = prepare_group(cs, n);
    if (rc) return rc;
    for (int r = 0; r < n; r++) cs[r]->dist->prepared = true;
    return NPRSPH_OK;
}

// ev (nullable): 5 events recorded on rank 0's stream at the stage boundaries
// [prepare | rho | halo(v,rho) | force | integrate]
int step_group(nprsph_ctx** cs, int n, cudaEvent_t* ev = nullptr) {
    Xfer x[64] = {};
    if (ev) CK(cs[0], cudaEventRecord(ev[0], cs[0]->stream));
    int rc = ensure_prepared(cs, n);
    if (rc) return rc;
    if (ev) CK(cs[0], cudaEventRecord(ev[1], cs[0]->stream));
    for (int r = 0; r < n; r++) {
        nprsph_ctx* c = cs[r]; DistState* d = c->dist;
        CK(c, cudaSetDevice(c->cfg.device));
        CK(c, cudaEventRecord(d->ev_work0, c->stream));
        launch_rho(c->pos[c->cur], c->vel[c->cur], nullptr, c->cell_start, d->own_off, d->n_own, d->lg,
                   c->sph, nullptr, c->hitmask, d->cap_total, c->stream);
        CK(c, cudaEventRecord(d->ev_work1, c->stream));
        d->work_timed = true;
        float4* v = c->vel[c->cur];
        const uint32_t right0 = d->own_off + d->n_own - d->n_nan - d->hR;
        x[r] = {v + d->own_off, (d->rank > 0 ? d->hL : 0u) * sizeof(float4),
                v + right0, (d->rank + 1 < d->world ? d->hR : 0u) * sizeof(float4),
                v + d->own_off - d->gL, d->gL * sizeof(float4),
                v + d->own_off + d->n_own, d->gR * sizeof(float4)};
    }
    if (ev) CK(cs[0], cudaEventRecord(ev[2], cs[0]->stream));
    // (velocity, rho) of the boundary layers travels on the communication streams while the
    // compute stream runs the force pass for the interior slots, whose walks touch no ghost cell;
    // the two boundary layers (and the NaN block) follow once the ghosts' (v, rho) have landed.
    // Range edges are even slots so that the slot pairs of k_rho / k_force_records stay aligned.
    for (int r = 0; r < n; r++) CK(cs[r], cudaEventRecord(cs[r]->dist->ev_rho, cs[r]->stream));
    for (int r = 0; r < n; r++)                    // LOCAL: ranks share one compute stream, so the
        CK(cs[r], cudaStreamWaitEvent(cs[r]->dist->comm_stream, cs[n - 1]->dist->ev_rho, 0));   // last event covers all
    rc = exchange_group(cs, n, x, true);
    if (rc) return rc;
    for (int r = 0; r < n; r++) CK(cs[r], cudaEventRecord(cs[r]->dist->ev_halo, cs[r]->dist->comm_stream));
    if (ev) CK(cs[0], cudaEventRecord(ev[3], cs[0]->stream));
    for (int r = 0; r < n; r++) {
        nprsph_ctx* c = cs[r]; DistState* d = c->dist;
        CK(c, cudaSetDevice(c->cfg.device));
        const uint32_t own_end = d->own_off + d->n_own;
        const uint32_t right0 = own_end - d->n_nan - d->hR;
        uint32_t in0 = (d->rank > 0) ? ((d->own_off + d->hL + 1u) & ~1u) : d->own_off;
        uint32_t in1 = (d->rank + 1 < d->world) ? (right0 & ~1u) : own_end;
        if (in0 > own_end) in0 = own_end;
        if (in1 < in0) in1 = in0;
        const float4* P = c->pos[c->cur]; const float4* V = c->vel[c->cur];
        // Passes 2 + 3 as one launch per slot range when the density pass left its column records
        // (the single-GPU step's fused kernel, plus slab keys and the next step's classification in
        // its epilogue); otherwise force launches followed by k_integrate_slab.
        const bool fused = !(c->cfg.flags & NPRSPH_FLAG_NO_FUSE) && c->hitmask && records_fit(d->lg.reach, d->cap_total);
        const int nxt = 1 - c->cur;
        const SlabNext sn = slab_next(c, true);
        if (fused) {
            // The interior range runs on the compute stream.  The two boundary ranges are small grids
            // whose threads still pay the latency of a whole walk (60-90 us each): queued behind the
            // interior they added 0.18 ms to the lattice step.  They go to the communication stream,
            // behind the (v, rho) halo that stream carries, and run concurrently with the interior
            // range as soon as the ghosts have landed.  All three share one queue of deferred slots,
            // drained on the compute stream once the boundary ranges are done.
            rc = reset_counts(c);
            if (rc) return rc;
            launch_force_queue_reset(c->hitmask, d->cap_total, d->lg.reach, c->stream);
            CK(c, cudaEventRecord(d->ev_reset, c->stream));
            auto fused_range = [&](uint32_t first, uint32_t count, cudaStream_t st) {
                launch_force_integrate_slab(P, V, c->frc[0], c->cell_start, first, count, d->lg, c->sph, c->hitmask,
                                            d->cap_total, c->pos[nxt], c->vel[nxt], c->keys[0], d->own_off,
                                            c->colliders, sn, st);
            };
            fused_range(in0, in1 - in0, c->stream);
            cudaStream_t bs = d->comm_stream;
            CK(c, cudaStreamWaitEvent(bs, d->ev_reset, 0));
            for (int q = 0; q < n; q++)              // LOCAL: a neighbour's receive may sit on ITS stream
                if (q != r) CK(c, cudaStreamWaitEvent(bs, cs[q]->dist->ev_halo, 0));
            fused_range(d->own_off, in0 - d->own_off, bs);
            fused_range(in1, own_end - in1, bs);
            CK(c, cudaEventRecord(d->ev_bdone, bs));
            CK(c, cudaStreamWaitEvent(c->stream, d->ev_bdone, 0));
            launch_force_deferred_slab(P, V, c->frc[0], c->cell_start, d->lg, c->sph, c->hitmask, d->cap_total,
                                       c->pos[nxt], c->vel[nxt], c->keys[0], d->own_off, c->colliders, sn, c->stream);
            if (ev && r == 0) CK(c, cudaEventRecord(ev[4], c->stream));
            c->cur = nxt;
        } else {
            auto force_range = [&](uint32_t first, uint32_t count) {
                launch_force(P, V, c->frc[0], c->cell_start, first, count, d->lg, c->sph, nullptr, c->hitmask,
                             d->cap_total, c->stream);
            };
            force_range(in0, in1 - in0);
            for (int q = 0; q < n; q++) CK(c, cudaStreamWaitEvent(c->stream, cs[q]->dist->ev_halo, 0));
            force_range(d->own_off, in0 - d->own_off);
            force_range(in1, own_end - in1);
            if (ev && r == 0) CK(c, cudaEventRecord(ev[4], c->stream));
            rc = reset_counts(c);
            if (rc) return rc;
            if (d->n_own)
                k_integrate_slab<<<blocks_for(d->n_own), TPB, 0, c->stream>>>(
                    c->pos[c->cur] + d->own_off, c->vel[c->cur] + d->own_off, c->frc[0] + d->own_off,
                    c->keys[0], d->n_own, c->sph, c->colliders, sn);
        }
        d->classified = true;                       // counts and migration buffers are ready
        CK(c, cudaGetLastError());
        d->prepared = false;                        // positions moved: keys/ghosts/table are stale
        d->steps_done++;
        c->steps_done++;
    }
    return NPRSPH_OK;
}

int check_group(nprsph_ctx** cs, int n) {
    if (!cs || n < 1 || n > 64) return NPRSPH_ERR_INVALID;
    for (int r = 0; r < n; r++) {
        if (!cs[r]) return NPRSPH_ERR_INVALID;
        if (cs[r]->sticky) return cs[r]->sticky;
        if (!cs[r]->dist) return fail(cs[r], NPRSPH_ERR_STATE, "nprsph_dist_init() was not called%s");
        if (n > 1 && (cs[r]->dist->transport != NPRSPH_TRANSPORT_LOCAL || cs[r]->dist->rank != r ||
                      cs[r]->dist->world != n || cs[r]->stream != cs[0]->stream))
            return fail(cs[r], NPRSPH_ERR_INVALID,
                        "several contexts in one call need the LOCAL transport, ranks 0..n-1 and one shared stream%s");
    }
    return NPRSPH_OK;
}

}  // namespace

int nprsph::slab_world(const nprsph_ctx* c) { return c->dist ? c->dist->world : 1; }

void nprsph::dist_destroy(nprsph_ctx* c) {
    DistState* d = c->dist;
    if (!d) return;
    if (d->nccl_comm && nccl() && nccl()->CommDestroy) nccl()->CommDestroy((ncclComm_t)d->nccl_comm);
    cudaFree(d->sendL); cudaFree(d->sendR); cudaFree(d->recv);
    cudaFree(d->d_counts); cudaFree(d->mig_ids); cudaFree(d->mig_sort_ws);
    if (d->h_counts) cudaFreeHost(d->h_counts);
    if (d->comm_stream) { cudaStreamSynchronize(d->comm_stream); cudaStreamDestroy(d->comm_stream); }
    if (d->ev_rho) cudaEventDestroy(d->ev_rho);
    if (d->ev_reset) cudaEventDestroy(d->ev_reset);
    if (d->ev_bdone) cudaEventDestroy(d->ev_bdone);
    if (d->ev_work0) cudaEventDestroy(d->ev_work0);
    if (d->ev_work1) cudaEventDestroy(d->ev_work1);
    if (d->ev_halo) cudaEventDestroy(d->ev_halo);
    delete d;
    c->dist = nullptr;
}

// ================================ C ABI ==========================================================
extern "C" {

int nprsph_slab_partition(const uint64_t* hist, int dimx, int world, int min_width, int32_t* bounds) {
    if (!hist || !bounds || dimx < 1 || world < 1 || min_width < 1 || (int64_t)world * min_width > dimx)
        return NPRSPH_ERR_INVALID;
    uint64_t total = 0;
    for (int x = 0; x < dimx; x++) total += hist[x];
    bounds[0] = 0;
    bounds[world] = dimx;
    uint64_t acc = 0;
    int x = 0;
    for (int r = 1; r < world; r++) {
        // smallest boundary whose prefix reaches r/world of the particles, leaving every slab
        // (also the remaining ones) at least min_width cells
        const uint64_t target = (total * (uint64_t)r + (uint64_t)world - 1) / (uint64_t)world;
        const int lo = bounds[r - 1] + min_width, hi = dimx - (world - r) * min_width;
        while (x < hi && (acc < target || x < lo)) acc += hist[x++];
        while (x > hi) acc -= hist[--x];
        bounds[r] = x;
    }
    return NPRSPH_OK;
}

static_assert(NPRSPH_SLAB_COUNTER_WORDS == CNT_WORDS && NPRSPH_CNT_OWN == CNT_NOWN && NPRSPH_CNT_FREE == CNT_FREE &&
              NPRSPH_CNT_WIDTH == CNT_WIDTH && NPRSPH_CNT_CAP_MIGRATE == CNT_CAPMIG && NPRSPH_CNT_HALO_R == CNT_HALO_R &&
              NPRSPH_CNT_COST_US == CNT_COST,
              "public counter layout");
int nprsph_slab_face_move(const uint32_t* a, const uint32_t* b, int reach, uint32_t cap_ghost, int by_time) {
    if (!a || !b || reach < 1) return 0;
    return face_move(a, b, reach, cap_ghost, by_time != 0);
}

int nprsph_dist_unique_id(uint8_t id[128]) {
    if (!id) return NPRSPH_ERR_INVALID;
    NcclApi* api = nccl();
    if (!api) return NPRSPH_ERR_UNSUPPORTED;
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != ncclSuccess) return NPRSPH_ERR_COMM;
    static_assert(sizeof u == 128, "ncclUniqueId size");
    memcpy(id, &u, 128);
    return NPRSPH_OK;
}

int nprsph_dist_init(nprsph_ctx* c, const nprsph_dist_config* cfg) {
    GUARD(c);
    if (!cfg || cfg->struct_size != sizeof(nprsph_dist_config) || cfg->world < 1 || cfg->rank < 0 ||
        cfg->rank >= cfg->world)
        return fail(c, NPRSPH_ERR_INVALID, "bad nprsph_dist_config%s");
    if (c->dist) return fail(c, NPRSPH_ERR_STATE, "nprsph_dist_init() called twice%s");
    DistState* d = new (std::nothrow) DistState();
    if (!d) return fail(c, NPRSPH_ERR_NOMEM, "out of host memory%s");
    d->rank = cfg->rank; d->world = cfg->world; d->transport = cfg->transport;
    c->dist = d;
    CK(c, cudaMalloc(&d->d_counts, (3 * CNT_WORDS + ERR_WORDS) * sizeof(uint32_t)));
    CK(c, cudaMemset(d->d_counts, 0, (3 * CNT_WORDS + ERR_WORDS) * sizeof(uint32_t)));
    CK(c, cudaMallocHost(&d->h_counts, (3 * CNT_WORDS + ERR_WORDS) * sizeof(uint32_t)));
    // (a decision sees the neighbours' counts of the step before: at least every second step)
    // rebalance_every < 0: balance the measured time of the density pass instead of particle counts
    d->balance_time = cfg->rebalance_every < 0;
    { const int every = cfg->rebalance_every < 0 ? -cfg->rebalance_every : cfg->rebalance_every;
      d->rebalance_every = every > 0 ? (every < 2 ? 2 : every) : 0; }
    CK(c, cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking));
    CK(c, cudaEventCreateWithFlags(&d->ev_rho, cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&d->ev_halo, cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&d->ev_reset, cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&d->ev_bdone, cudaEventDisableTiming));
    CK(c, cudaEventCreate(&d->ev_work0));
    CK(c, cudaEventCreate(&d->ev_work1));
    if (cfg->transport == NPRSPH_TRANSPORT_NCCL && cfg->world > 1) {
        NcclApi* api = nccl();
        if (!api) return fail(c, NPRSPH_ERR_UNSUPPORTED, "libnccl.so.2 not found%s");
        ncclUniqueId u;
        memcpy(&u, cfg->nccl_id, 128);
        ncclComm_t comm;
        NCK(c, api->CommInitRank(&comm, cfg->world, u, cfg->rank));
        d->nccl_comm = comm;
    } else if (cfg->transport != NPRSPH_TRANSPORT_NCCL && cfg->transport != NPRSPH_TRANSPORT_LOCAL) {
        return fail(c, NPRSPH_ERR_INVALID, "unknown transport%s");
    }
    // capacities requested by the caller (0 = derived from the scene)
    d->cap_own = (uint32_t)cfg->max_own; d->cap_ghost = (uint32_t)cfg->max_ghost; d->cap_mig = (uint32_t)cfg->max_migrate;
    // the single-GPU particle buffer of create() is not used in slab mode
    c->n = 0;
    return NPRSPH_OK;
}

int nprsph_dist_scene_block(nprsph_ctx* c, int nx, int ny, int nz, float spacing, const float origin[3],
                            float jitter, uint32_t seed) {
    GUARD(c);
    DistState* d = c->dist;
    if (!d) return fail(c, NPRSPH_ERR_STATE, "nprsph_dist_init() was not called%s");
    if (nx < 1 || ny < 1 || nz < 1 || !(spacing > 0.0f) || (uint64_t)nx * ny * nz >= (1ull << 32))
        return fail(c, NPRSPH_ERR_INVALID, "bad block%s");
    int rc = refresh_params(c);
    if (rc) return rc;
    // every rank generates only the lattice planes within one cell of its slab: a jitter of a cell or
    // more could move a particle into a slab whose rank never generated it
    if (!(jitter < c->cell_size)) return fail(c, NPRSPH_ERR_INVALID, "slab scene: jitter must stay below one grid cell%s");
    const float o[3] = {origin ? origin[0] : 0.f, origin ? origin[1] : 0.f, origin ? origin[2] : 0.f};
    // every rank derives the same count-balanced slab boundaries from the lattice planes
    const GridDev& g = c->grid;
    const int R = g.reach;
    std::string herr;
    uint64_t* hist = new (std::nothrow) uint64_t[g.dim[0]]();
    int32_t* bounds = new (std::nothrow) int32_t[d->world + 1];
    if (!hist || !bounds) { delete[] hist; delete[] bounds; return fail(c, NPRSPH_ERR_NOMEM, "out of host memory%s"); }
    auto plane_cell = [&](int i) {
        float u = ((float)i * spacing + o[0] - g.lo[0]) * g.inv_cell;
        if (!(u >= 0.0f)) u = 0.0f;
        if (u > (float)(g.dim[0] - 1)) u = (float)(g.dim[0] - 1);
        return (int)u;
    };
    for (int i = 0; i < nx; i++) hist[plane_cell(i)] += (uint64_t)ny * nz;
    rc = nprsph_slab_partition(hist, g.dim[0], d->world, 2 * R, bounds);
    if (rc) { delete[] hist; delete[] bounds; return fail(c, NPRSPH_ERR_INVALID, "grid too narrow for this many ranks (each slab needs 2*reach cells)%s"); }
    d->X0 = bounds[d->rank]; d->X1 = bounds[d->rank + 1]; d->W = d->X1 - d->X0;
    d->X0_next = d->X0; d->X1_next = d->X1;
    // candidate lattice planes: those whose cell (ignoring jitter) is within one cell of the slab
    int i0 = nx, i1 = 0;
    for (int i = 0; i < nx; i++) {
        const int cx = plane_cell(i);
        if (cx >= d->X0 - 1 && cx <= d->X1) { if (i < i0) i0 = i; if (i + 1 > i1) i1 = i + 1; }
    }
    uint64_t own_est = 0;
    for (int x = d->X0; x < d->X1; x++) own_est += hist[x];
    delete[] hist; delete[] bounds;
    const int ni = i1 > i0 ? i1 - i0 : 0;
    const uint64_t n_cand = (uint64_t)ni * ny * nz;
    const uint64_t n_global = (uint64_t)nx * ny * nz;
    // capacities: own = 1.5x the fair share (the static slabs let fluid pile up downstream),
    // ghosts = the particles of `reach`+2 cell layers of the block's cross-section, twice over
    const double layers = (double)(R + 2) / ((double)spacing * g.inv_cell) + 2.0;
    uint64_t cap_ghost = d->cap_ghost ? d->cap_ghost : (uint64_t)(2.0 * layers * ny * nz) + 4096;
    uint64_t cap_own = d->cap_own ? d->cap_own : (uint64_t)(1.5 * (double)n_global / d->world) + 65536;
    if (cap_own < n_cand) cap_own = n_cand;
    if (cap_own < own_est) cap_own = own_est;
    uint64_t cap_mig = d->cap_mig ? d->cap_mig : cap_ghost / 2 + 1024;
    rc = alloc_slab(c, cap_own, cap_ghost, cap_mig);
    if (rc) return rc;
    rc = setup_local_grid(c);
    if (rc) return rc;
    c->cur = 0;
    if (n_cand)
        k_slab_scene<<<blocks_for(n_cand), TPB, 0, c->stream>>>(
            c->pos[0] + d->own_off, c->vel[0] + d->own_off, c->keys[0], i0, ni, ny, nz, spacing, o[0],
            o[1], o[2], jitter, seed, d->lg, d->W, d->R);
    CK(c, cudaGetLastError());
    d->n_own = (uint32_t)n_cand;
    d->first_prepare = true;
    d->xmax_known = false;
    d->gather_force = false;         // (the scene's force array is zero)
    d->prepared = false;
    d->classified = false;
    d->ready = true;
    d->steps_done = 0;
    d->migrated_total = 0;
    d->rebalanced = 0;
    CK(c, cudaMemsetAsync(c->frc[0], 0, (size_t)d->cap_total * sizeof(float4), c->stream));
    return NPRSPH_OK;
}

int nprsph_dist_link_local(nprsph_ctx** ranks, int n) {
    if (!ranks || n < 1) return NPRSPH_ERR_INVALID;
    for (int r = 0; r < n; r++) {
        if (!ranks[r] || !ranks[r]->dist) return NPRSPH_ERR_STATE;
        ranks[r]->dist->left = r > 0 ? ranks[r - 1] : nullptr;
        ranks[r]->dist->right = r + 1 < n ? ranks[r + 1] : nullptr;
    }
    return check_group(ranks, n);
}

int nprsph_dist_step(nprsph_ctx** ranks, int n_local, int steps) {
    int rc = check_group(ranks, n_local);
    if (rc) return rc;
    if (steps < 0) return NPRSPH_ERR_INVALID;
    for (int r = 0; r < n_local; r++) {
        if (!ranks[r]->dist->ready) return fail(ranks[r], NPRSPH_ERR_STATE, "no scene distributed yet%s");
 
        const bool dirty = ranks[r]->params_dirty;
        rc = refresh_params(ranks[r]);
        if (rc) return rc;
        if (dirty) {                                 // constants may change freely; the grid may not
            const GridDev old = ranks[r]->dist->lg;
            rc = setup_local_grid(ranks[r]);
            if (rc) return rc;
            if (memcmp(&old, &ranks[r]->dist->lg, sizeof old) != 0)
                return fail(ranks[r], NPRSPH_ERR_UNSUPPORTED,
                            "changing the smoothing length or the box after the scene was distributed is not supported%s");
        }
    }
    if (ranks[0]->paused) return NPRSPH_OK;                  // if (simulate) ..., Main.cpp:293
    for (int s = 0; s < steps; s++) {
        rc = step_group(ranks, n_local);
        if (rc) return rc;
    }
    return NPRSPH_OK;
}

int nprsph_dist_download(nprsph_ctx** ranks, int n_local, int which, nprsph_particle* records,
                         uint32_t* ids, uint64_t capacity, uint64_t* n_out) {
    int rc = check_group(ranks, n_local);
    if (rc) return rc;
    if (which < 0 || which >= n_local || !n_out) return NPRSPH_ERR_INVALID;
    // After a step every particle is still held by exactly one rank (a leaver sits in its old
    // rank's slots until the next prepare hands it over), and position, velocity, force and density
    // share one slot order: the records are packed as they are.  Only a freshly distributed scene
    // still holds candidates of other slabs and needs the (collective) prepare first.
    bool fresh = false;
    for (int r = 0; r < n_local; r++) fresh = fresh || ranks[r]->dist->first_prepare;
    if (fresh) {
        rc = ensure_prepared(ranks, n_local);
        if (rc) return rc;
    }
    nprsph_ctx* c = ranks[which]; DistState* d = c->dist;
    CK(c, cudaSetDevice(c->cfg.device));
    *n_out = d->n_own;
    if (!records && !ids) return NPRSPH_OK;
    if (capacity < d->n_own || !records || !ids) return fail(c, NPRSPH_ERR_INVALID, "download buffers too small%s");
    if (d->n_own) {
        uint32_t* d_ids = c->vals[1];
        k_pack_records<<<blocks_for(d->n_own), TPB, 0, c->stream>>>(
            c->pos[c->cur] + d->own_off, c->vel[c->cur] + d->own_off, c->frc[0] + d->own_off, d->n_own,
            (float4*)c->aos, d_ids);
        CK(c, cudaMemcpyAsync(records, c->aos, (size_t)d->n_own * 64, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaMemcpyAsync(ids, d_ids, (size_t)d->n_own * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    return NPRSPH_OK;
}

int nprsph_dist_upload(nprsph_ctx* c, const nprsph_particle* records, const uint32_t* ids, uint64_t n) {
    GUARD(c);
    DistState* d = c->dist;
    if (!d || !d->ready) return fail(c, NPRSPH_ERR_STATE, "distribute a scene first (slab geometry and capacities)%s");
    if (n > d->cap_own || (n && (!records || !ids))) return fail(c, NPRSPH_ERR_INVALID, "bad upload%s");
    const SlabNext sn = slab_next(c, false);
    if (n) {
        uint32_t* d_ids = c->vals[1];
        CK(c, cudaMemcpyAsync(c->aos, records, n * 64, cudaMemcpyHostToDevice, c->stream));
        CK(c, cudaMemcpyAsync(d_ids, ids, n * 4, cudaMemcpyHostToDevice, c->stream));
        k_unpack_records<<<blocks_for(n), TPB, 0, c->stream>>>(
            (const float4*)c->aos, d_ids, (uint32_t)n, c->pos[c->cur] + d->own_off,
            c->vel[c->cur] + d->own_off, c->frc[0] + d->own_off, c->keys[0], sn.g, sn.W, sn.R);
        CK(c, cudaGetLastError());
    }
    d->n_own = (uint32_t)n;
    d->xmax_known = false;
    d->first_prepare = false;        // a record of the slab next door is handed over by the next prepare
    d->gather_force = true;          // the uploaded force / pressure columns follow the re-sort
    d->prepared = false;
    d->classified = false;
    return NPRSPH_OK;
}

// A host that re-uploads the SAME particle lists step after step (a replay, the e2e measurement) must
// keep the slab faces where they are: every face move shifts one more x layer of those lists into the
// slab next door, and an uploaded particle may lie at most `reach` layers beyond a face (it is handed
// over like a leaver, k_unpack_migrants).  Every rank of the group must make the same call (the two
// ranks of a face decide together).  A move already decided still takes effect at the next step.
int nprsph_dist_freeze_faces(nprsph_ctx* c, int frozen) {
    GUARD(c);
    if (!c->dist) return fail(c, NPRSPH_ERR_STATE, "not a slab rank (nprsph_dist_init)%s");
    c->dist->faces_frozen = frozen != 0;
    return NPRSPH_OK;
}

// The per-step traffic of a host application in slab mode: the working arrays already are
// (x, y, z, id bits) and (vx, vy, vz, rho), so no record packing -- 32 + 16 bytes per particle.  The
// streams, events and staging buffers are the context's (api.cu uses the same for the single-GPU pair).
static int ensure_staging_slab(nprsph_ctx* c) {
    DistState* d = c->dist;
    if (!c->h2d_stream) {
        CK(c, cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
        CK(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < nprsph_ctx::STAGE_CHUNKS; i++) CK(c, cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) CK(c, cudaEventCreateWithFlags(&c->ev_imported[i], cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_published, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
    }
    if (c->stage_cap < d->cap_own) {
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, cudaStreamSynchronize(c->h2d_stream));
        CK(c, cudaStreamSynchronize(c->d2h_stream));
        c->stage_cap = 0;
        for (int i = 0; i < 2; i++) CK(c, realloc_dev(c->stage_in[i], 2 * (size_t)d->cap_own));
        CK(c, realloc_dev(c->stage_pos, (size_t)d->cap_own));
        c->stage_cap = d->cap_own;
    }
    return NPRSPH_OK;
}

int nprsph_dist_upload_state(nprsph_ctx* c, const float* pos4, const float* vel4, uint64_t n) {
    GUARD(c);
    DistState* d = c->dist;
    if (!d || !d->ready) return fail(c, NPRSPH_ERR_STATE, "distribute a scene first (slab geometry and capacities)%s");
    if (n > d->cap_own || (n && (!pos4 || !vel4))) return fail(c, NPRSPH_ERR_INVALID, "bad upload%s");
    const SlabNext sn = slab_next(c, false);
    if (n) {
        int rc = ensure_staging_slab(c);
        if (rc) return rc;
        // the copy stream brings the chunks in -- while the previous step still computes -- and the
        // compute stream moves each into the own slots as it lands
        const int b = c->stage_cur;
        c->stage_cur ^= 1;
        float4* sp = c->stage_in[b];
        float4* sv = sp + d->cap_own;
        CK(c, cudaStreamWaitEvent(c->h2d_stream, c->ev_imported[b], 0));     // its last import has read the buffer
        const uint64_t per = (n + nprsph_ctx::STAGE_CHUNKS - 1) / nprsph_ctx::STAGE_CHUNKS;
        for (int k = 0; k < nprsph_ctx::STAGE_CHUNKS; k++) {
            const uint64_t first = (uint64_t)k * per;
            if (first >= n) break;
            const uint64_t cnt = n - first < per ? n - first : per;
            CK(c, cudaMemcpyAsync(sp + first, pos4 + 4 * first, cnt * sizeof(float4), cudaMemcpyHostToDevice, c->h2d_stream));
            CK(c, cudaMemcpyAsync(sv + first, vel4 + 4 * first, cnt * sizeof(float4), cudaMemcpyHostToDevice, c->h2d_stream));
            CK(c, cudaEventRecord(c->ev_chunk[k], c->h2d_stream));
            CK(c, cudaStreamWaitEvent(c->stream, c->ev_chunk[k], 0));
            k_import_state_slab<<<blocks_for(cnt), TPB, 0, c->stream>>>(
                sp + first, sv + first, (uint32_t)cnt, c->pos[c->cur] + d->own_off + first,
                c->vel[c->cur] + d->own_off + first, c->keys[0] + first, sn.g, sn.W, sn.R);
        }
        CK(c, cudaEventRecord(c->ev_imported[b], c->stream));
        CK(c, cudaGetLastError());
    }
    d->n_own = (uint32_t)n;
    d->xmax_known = false;
    d->first_prepare = false;        // a particle of the slab next door is handed over by the next prepare
    d->gather_force = false;         // force is an output of the next step
    d->prepared = false;
    d->classified = false;
    return NPRSPH_OK;
}

int nprsph_dist_download_positions(nprsph_ctx** ranks, int n_local, int which, float* pos4,
                                   uint64_t capacity, uint64_t* n_out, uint32_t flags) {
    int rc = check_group(ranks, n_local);
    if (rc) return rc;
    if (which < 0 || which >= n_local || !n_out) return NPRSPH_ERR_INVALID;
    bool fresh = false;              // (see nprsph_dist_download)
    for (int r = 0; r < n_local; r++) fresh = fresh || ranks[r]->dist->first_prepare;
    if (fresh) {
        rc = ensure_prepared(ranks, n_local);
        if (rc) return rc;
    }
    nprsph_ctx* c = ranks[which]; DistState* d = c->dist;
    CK(c, cudaSetDevice(c->cfg.device));
    *n_out = d->n_own;
    if (!pos4) return NPRSPH_OK;
    if (capacity < d->n_own) return fail(c, NPRSPH_ERR_INVALID, "download buffer too small%s");
    if (!d->n_own) return NPRSPH_OK;
    rc = ensure_staging_slab(c);
    if (rc) return rc;
    // a copy of the own positions, so that the next step may overwrite the slots while they travel
    if (c->d2h_pending) CK(c, cudaStreamWaitEvent(c->stream, c->ev_copied, 0));    // the previous copy has left the staging buffer
    CK(c, cudaMemcpyAsync(c->stage_pos, c->pos[c->cur] + d->own_off, (size_t)d->n_own * sizeof(float4),
                          cudaMemcpyDeviceToDevice, c->stream));
    CK(c, cudaEventRecord(c->ev_published, c->stream));
    CK(c, cudaStreamWaitEvent(c->d2h_stream, c->ev_published, 0));
    CK(c, cudaMemcpyAsync(pos4, c->stage_pos, (size_t)d->n_own * sizeof(float4), cudaMemcpyDeviceToHost, c->d2h_stream));
    CK(c, cudaEventRecord(c->ev_copied, c->d2h_stream));
    c->d2h_pending = true;
    if (!(flags & NPRSPH_DOWNLOAD_ASYNC)) {
        CK(c, cudaStreamSynchronize(c->d2h_stream));
        c->d2h_pending = false;
    }
    return NPRSPH_OK;
}

int nprsph_dist_profile_step(nprsph_ctx** ranks, int n_local, int steps, float* stage_ms) {
    int rc = check_group(ranks, n_local);
    if (rc) return rc;
    if (steps < 1 || !stage_ms) return NPRSPH_ERR_INVALID;
    nprsph_ctx* c0 = ranks[0];
    CK(c0, cudaSetDevice(c0->cfg.device));
    cudaEvent_t ev[6];
    for (int i = 0; i < 6; i++) CK(c0, cudaEventCreate(&ev[i]));
    double acc[NPRSPH_NUM_STAGES] = {0};
    for (int s = 0; s < steps && rc == NPRSPH_OK; s++) {
        rc = step_group(ranks, n_local, ev);
        if (rc) break;
        CK(c0, cudaEventRecord(ev[5], c0->stream));
        for (int r = 0; r < n_local; r++) CK(ranks[r], cudaStreamSynchronize(ranks[r]->stream));
        // prepare (classify..sort..halo..table) is reported as SORT, the (v, rho) halo as REORDER
        const int stage_of[5] = {NPRSPH_STAGE_SORT, NPRSPH_STAGE_RHO, NPRSPH_STAGE_REORDER,
                                 NPRSPH_STAGE_FORCE, NPRSPH_STAGE_INTEGRATE};
        for (int i = 0; i < 5; i++) {
            float ms = 0.f;
            CK(c0, cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            acc[stage_of[i]] += ms;
        }
    }
    for (int i = 0; i < 6; i++) cudaEventDestroy(ev[i]);
    for (int i = 0; i < NPRSPH_NUM_STAGES; i++) stage_ms[i] = (float)(acc[i] / steps);
    return rc;
}

int nprsph_dist_get_info(nprsph_ctx* c, nprsph_dist_info* out) {
    GUARD(c);
    if (!out || !c->dist) return NPRSPH_ERR_INVALID;
    DistState* d = c->dist;
    memset(out, 0, sizeof *out);
    out->rank = d->rank; out->world = d->world;
    out->x_begin = d->X0; out->x_end = d->X1;
    out->num_own = d->n_own; out->ghosts_left = d->gL; out->ghosts_right = d->gR;
    out->migrated_total = d->migrated_total; out->steps_done = d->steps_done;
    out->cap_own = d->cap_own; out->cap_ghost = d->cap_ghost;
    out->nan_particles = d->n_nan;
    out->rebalanced = d->rebalanced;
    out->last_migrated = d->last_migrated;
    out->sort_bits = (uint32_t)d->sort_bits;
    out->sort_passes = (uint32_t)sort_num_passes(d->sort_bits);
    return NPRSPH_OK;
}

}  // extern "C"
