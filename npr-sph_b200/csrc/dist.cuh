// dist.cuh -- state of the slab-decomposed (multi-GPU) step, one context per GPU / process.
#pragma once

#include "context.cuh"
#include "slab.cuh"

namespace nprsph {

struct DistState {
    int rank = 0, world = 1;
    int transport = 0;                 // NPRSPH_TRANSPORT_*
    void* nccl_comm = nullptr;
    cudaStream_t comm_stream = nullptr;   // halo traffic that overlaps the interior force pass
    cudaEvent_t ev_rho = nullptr, ev_halo = nullptr;
    cudaEvent_t ev_reset = nullptr, ev_bdone = nullptr;   // counters reset / boundary force ranges done
    nprsph_ctx* left = nullptr;        // local transport only
    nprsph_ctx* right = nullptr;

    // slab geometry (global x cell indices) and the local grid derived from ctx->grid
    int X0 = 0, X1 = 0, W = 0, R = 1;
    GridDev lg;
    int key_bits = 1;                  // bits of the largest cell key of the local grid (+ sentinels)
    int sort_bits = 1;                 // bits the last prepare actually sorted on (occupied x layers)
    bool ready = false;                // scene distributed
    bool prepared = false;             // ownership, ghosts and cell table match the current positions
    bool classified = false;           // d_counts / sendL / sendR already hold the classification of keys[0]
                                       // (k_integrate_slab did it); cleared by anything else that rewrites keys

    // slot layout: [ghost L | own | ghost R]; own particles start at slot own_off
    uint32_t cap_ghost = 0, cap_own = 0, cap_mig = 0, cap_total = 0, own_off = 0;
    uint32_t n_own = 0, n_nan = 0, gL = 0, gR = 0, hL = 0, hR = 0;
    uint32_t x_top = 0;                // highest local x cell layer any walk of this step can touch
    bool xmax_known = false;           // the sticky word STICKY_XMAX holds the top occupied layer of the last sort
    bool first_prepare = true;         // candidates that belong to other ranks are dropped, not sent
    uint64_t migrated_total = 0, steps_done = 0;
    uint32_t last_migrated = 0;        // particles this rank handed over in the last prepare

    Migrant *sendL = nullptr, *sendR = nullptr, *recv = nullptr;
    uint32_t* d_counts = nullptr;      // [3][CNT_WORDS]: mine, from left, from right; then [ERR_WORDS] sticky
    uint32_t* h_counts = nullptr;      // pinned mirror
    bool gather_force = false;         // the next prepare also permutes the force array (after an upload /
                                       // scene: later steps recompute the force before anyone reads it)
    // slab faces the NEXT prepare will use (re-balancing moves a face by one x layer; the step in
    // between computes its keys against them)
    int X0_next = 0, X1_next = 0;
    int rebalance_every = 0;           // steps between re-balancing decisions (0 = static slabs)
    bool balance_time = false;         // balance the measured density-pass time instead of particle counts
    cudaEvent_t ev_work0 = nullptr, ev_work1 = nullptr;   // around the density pass of the last step
    bool work_timed = false;
    float cost_ms = 0.f;               // smoothed duration of the density pass (the rank's work per step)
    uint64_t rebalanced = 0;           // face moves so far (both faces of this rank)
    bool faces_frozen = false;         // nprsph_dist_freeze_faces: no further re-balancing decisions
    uint32_t* mig_ids = nullptr;       // [4][2*cap_mig] scratch: ids, iota, sorted ids, order
    void* mig_sort_ws = nullptr;
};

}  // namespace nprsph
