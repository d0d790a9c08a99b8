// context.cuh -- the context object behind nprsph_ctx* and the error helpers shared by the host
// translation units (api.cu: single-GPU runtime + C ABI, dist.cu: slab-decomposed multi-GPU step).
#pragma once

#include "../../include/nprsph.h"

#include <stdio.h>

#include <string>

#include "kernels.cuh"
#include "sort.cuh"

namespace nprsph { struct DistState; }

struct nprsph_ctx {
    nprsph_config cfg;
    nprsph_constants consts;
    nprsph_boundary bounds;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool paused = true;                 // `bool simulate;` starts false, Main.cpp:87
    std::string err;
    int sticky = 0;

    struct { int nx, ny, nz; float spacing, origin[3], jitter; uint32_t seed; } scene;

    uint64_t n = 0, cap = 0;
    void* aos = nullptr;                // Particle[n], original order (SSBO binding 0)
    float4* pos[2] = {nullptr, nullptr};
    float4* vel[2] = {nullptr, nullptr};
    float4* frc[2] = {nullptr, nullptr};
    int cur = 0;
    nprsph::ColliderSet colliders = {};         // static obstacles of the integrate pass (nprsph_set_colliders)
    uint32_t* keys[2] = {nullptr, nullptr};
    uint32_t* vals[2] = {nullptr, nullptr};
    uint32_t* sorted_keys = nullptr;
    uint32_t* last_perm = nullptr;
    uint32_t* counts_rho = nullptr;
    uint32_t* counts_force = nullptr;
    void* sort_ws = nullptr;
    uint32_t* hitmask = nullptr;        // column records, rho -> force (sph_passes.cu); rec_buffer_words(cap)
    bool mask_valid = false;

    uint32_t* cell_start = nullptr;
    size_t cell_cap = 0;
    uint4* gap_list = nullptr;
    size_t gap_cap = 0;
    uint32_t* gap_count = nullptr;      // also scratch for the NaN counter (8 bytes)

    nprsph::GridDev grid;
    nprsph::SphDev sph;
    float cell_size = 0.f;
    int key_bits = 1;
    bool params_dirty = true;
    bool keys_valid = false;
    bool grid_valid = false;
    bool aos_stale = false;             // SoA state is newer than the AoS view
    uint64_t steps_done = 0;

    cudaGraphicsResource* gl_res = nullptr;

    nprsph::DistState* dist = nullptr;   // non-null once nprsph_dist_init() succeeded
};

namespace nprsph {

// error bookkeeping: message into the context (or the thread-local create-error slot), CUDA
// errors become sticky
int fail(nprsph_ctx* c, int code, const char* fmt, const char* detail = "");

#define CK(ctx, call)                                                                     \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) return nprsph::fail((ctx), NPRSPH_ERR_CUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

#define GUARD(ctx)                                                                        \
    do {                                                                                  \
        if (!(ctx)) return NPRSPH_ERR_INVALID;                                            \
        if ((ctx)->sticky) return (ctx)->sticky;                                          \
        cudaError_t e_ = cudaSetDevice((ctx)->cfg.device);                                \
        if (e_ != cudaSuccess) return nprsph::fail((ctx), NPRSPH_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e_)); \
    } while (0)

// entry points of the single-context path are not valid once the context runs in slab mode
#define SINGLE_ONLY(ctx)                                                                  \
    do {                                                                                  \
        if ((ctx)->dist) return nprsph::fail((ctx), NPRSPH_ERR_STATE,                     \
            "this context is a slab rank (nprsph_dist_init): use the nprsph_dist_* entry points%s"); \
    } while (0)

// api.cu
int refresh_params(nprsph_ctx* c);
template <typename T>
cudaError_t realloc_dev(T*& p, size_t count) {
    if (p) { cudaError_t e = cudaFree(p); p = nullptr; if (e != cudaSuccess) return e; }
    return count ? cudaMalloc(&p, count * sizeof(T)) : cudaSuccess;
}

// dist.cu
void dist_destroy(nprsph_ctx* c);
int slab_world(const nprsph_ctx* c);      // number of slab ranks (1 without nprsph_dist_init)

}  // namespace nprsph
