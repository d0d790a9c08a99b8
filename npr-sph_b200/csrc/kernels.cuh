// kernels.cuh -- launchers of every device kernel (implemented in grid.cu / sph_passes.cu).
#pragma once
#include "common.cuh"

namespace nprsph {

// grid.cu
void launch_import(const void* aos, float4* posid, float4* velrho, float4* forcep, uint32_t n,
                   cudaStream_t st);
void launch_publish(const float4* posid, const float4* velrho, const float4* forcep, void* aos,
                    uint32_t n, cudaStream_t st);
// positions only, original order (out[id] = x, y, z, the record's own w lane)
void launch_publish_positions(const float4* posid, const void* aos, float4* out, uint32_t n, cudaStream_t st);
// host state chunk [first, first + count) (positions, velocities; original order) -> SoA
void launch_import_state(const float4* pos_in, const float4* vel_in, uint32_t first, uint32_t count,
                         float4* posid, float4* velrho, float4* forcep, cudaStream_t st);
void launch_keys(const float4* posid, uint32_t* keys, uint32_t n, const GridDev& g, cudaStream_t st);
size_t gap_list_capacity(uint32_t num_cells, uint64_t n);
void launch_reorder_cells(const uint32_t* sorted_keys, const uint32_t* perm, const float4* pos_in,
                          const float4* vel_in, const float4* force_in, float4* pos_out,
                          float4* vel_out, float4* force_out, uint32_t* cell_start,
                          uint32_t num_cells, uint32_t n, uint4* gap_list, uint32_t* gap_count,
                          bool with_force, int num_sms, cudaStream_t st);
void launch_fill_gaps(const uint4* gap_list, const uint32_t* gap_count, uint32_t* cell_start,
                      int num_sms, cudaStream_t st);
void launch_gather(const uint32_t* src_of_slot, const float4* pos_in, const float4* vel_in,
                   const float4* frc_in, float4* pos_out, float4* vel_out, float4* frc_out, uint32_t n,
                   cudaStream_t st);
void launch_count_nan(const float4* posid, uint32_t n, unsigned long long* out, cudaStream_t st);
void launch_slot_ids(const float4* posid, uint32_t* ids, uint32_t n, cudaStream_t st);

// sph_passes.cu
// whether the column records (rho -> force) can describe walks of this reach over `stride` slots
bool records_fit(int reach, uint64_t stride);
void launch_rho(const float4* posid, float4* velrho, float4* forcep_or_null,
                const uint32_t* cell_start, uint32_t first, uint32_t n, const GridDev& g, const SphDev& sp,
                uint32_t* counts_by_id, uint32_t* hitmask_or_null, uint32_t mask_stride,
                cudaStream_t st);
void launch_force(const float4* posid, const float4* velrho, float4* forcep,
                  const uint32_t* cell_start, uint32_t first, uint32_t n, const GridDev& g, const SphDev& sp,
                  uint32_t* counts_by_id, const uint32_t* hitmask_or_null, uint32_t mask_stride,
                  cudaStream_t st);
// Passes 2 + 3 in one launch (needs the density pass's column records): forces into forcep, the
// integrated state into the other buffer pair, next-step keys.  false: not applicable, nothing ran.
bool launch_force_integrate(const float4* posid, const float4* velrho, float4* forcep,
                            const uint32_t* cell_start, uint32_t n, const GridDev& g, const SphDev& sp,
                            uint32_t* counts_by_id, const uint32_t* records_or_null, uint32_t rec_stride,
                            float4* pos_next, float4* vel_next, uint32_t* keys_next,
                            const ColliderSet& cs, cudaStream_t st);
// measurement only: out5 += {distance tests, non-empty columns, pair walks, single walks, neighbours}
void launch_walk_stats(const float4* posid, const uint32_t* cell_start, uint32_t first, uint32_t n,
                       const GridDev& g, const SphDev& sp, unsigned long long* out5, cudaStream_t st);
struct SlabNext;
void launch_force_integrate_slab(const float4* posid, const float4* velrho, float4* forcep,
                                 const uint32_t* cell_start, uint32_t first, uint32_t n, const GridDev& g,
                                 const SphDev& sp, const uint32_t* records, uint32_t rec_stride,
                                 float4* pos_next, float4* vel_next, uint32_t* keys_next, uint32_t key_base,
                                 const ColliderSet& cs, const SlabNext& sn, cudaStream_t st);
void launch_force_queue_reset(const uint32_t* records, uint32_t rec_stride, int reach, cudaStream_t st);
void launch_force_deferred_slab(const float4* posid, const float4* velrho, float4* forcep,
                                const uint32_t* cell_start, const GridDev& g, const SphDev& sp,
                                const uint32_t* records, uint32_t rec_stride, float4* pos_next, float4* vel_next,
                                uint32_t* keys_next, uint32_t key_base, const ColliderSet& cs, const SlabNext& sn,
                                cudaStream_t st);
void launch_integrate(float4* posid, float4* velrho, const float4* forcep, uint32_t* keys,
                      uint32_t n, const GridDev& g, const SphDev& sp, const ColliderSet& cs,
                      cudaStream_t st);

}  // namespace nprsph
