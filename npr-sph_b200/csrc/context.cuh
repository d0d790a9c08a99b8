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
    uint32_t* hitmask = nullptr;        // column records, rho -> force (sph_passes.cu)
    size_t hitmask_words = 0;           // allocated: rec_buffer_words(record capacity, rec_cols_of(reach))
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
    bool gl_seeded = false;             // the registered GL buffer holds every lane of the records

    // streaming interface (nprsph_upload_state / nprsph_download_positions): two copy streams so
    // that a step's result leaves while the next step's inputs arrive and the step itself runs
    static constexpr int STAGE_CHUNKS = 4;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_chunk[STAGE_CHUNKS] = {}, ev_imported[2] = {}, ev_published = nullptr, ev_copied = nullptr;
    float4* stage_in[2] = {nullptr, nullptr};   // [pos | vel] of one upload each, double-buffered
    float4* stage_pos = nullptr;                // positions in original order for the D2H copy
    uint64_t stage_cap = 0;
    int stage_cur = 0;
    bool d2h_pending = false;

    // nprsph_step as one CUDA graph launch (api.cu: step_any).  A step is a fixed sequence of ~15
    // launches whose arguments only change when the caller changes something; StepSig is everything
    // those arguments are made of, StepPost the host-side state a step leaves behind.
    struct StepSig {
        uint64_t n, cap;
        int cur, key_bits, num_sms;
        uint32_t flags;
        int keys_valid, grid_valid, mask_valid;
        size_t hitmask_words;
        const void* ptr[18];
        nprsph::GridDev grid;
        nprsph::SphDev sph;
        nprsph::ColliderSet colliders;
    };
    struct StepPost { int cur; bool keys_valid, grid_valid, mask_valid, aos_stale; uint32_t* sorted_keys; uint32_t* last_perm; };
    cudaGraphExec_t step_graph = nullptr;
    StepSig graph_sig, plain_sig;       // launch parameters of the recorded step / of the last plain step
    StepPost graph_post;
    bool plain_sig_valid = false;
    bool graph_off = false;             // NPRSPH_FLAG_NO_GRAPH, NPRSPH_NO_GRAPH=1, or a capture that failed
    uint64_t graph_steps = 0;

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
// (re)allocate the column-record buffer for `slots` slots and walks of `reach` cells; frees it when
// the records cannot describe such walks (the passes then re-test their candidates)
int ensure_records(nprsph_ctx* c, uint64_t slots, int reach);
int effective_subdiv(const nprsph_ctx* c);
template <typename T>
cudaError_t realloc_dev(T*& p, size_t count) {
    if (p) { cudaError_t e = cudaFree(p); p = nullptr; if (e != cudaSuccess) return e; }
    return count ? cudaMalloc(&p, count * sizeof(T)) : cudaSuccess;
}

// dist.cu
void dist_destroy(nprsph_ctx* c);
int slab_world(const nprsph_ctx* c);      // number of slab ranks (1 without nprsph_dist_init)

}  // namespace nprsph
