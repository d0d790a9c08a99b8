// sort.cuh -- interface of the hand-written onesweep radix sort (onesweep.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace nprsph {

constexpr int SORT_MAX_RADIX = 512;                    // 8-bit digits, or 9-bit when that saves a pass
#ifndef NPRSPH_SORT_THREADS
#define NPRSPH_SORT_THREADS 256
#endif
constexpr int SORT_THREADS = NPRSPH_SORT_THREADS;      // 256 (16 pairs per thread) or 512 (8 pairs per thread)
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_TILE = 4096;                        // pairs per CTA
constexpr int SORT_ITEMS = SORT_TILE / SORT_THREADS;
constexpr int SORT_MAX_PASSES = 4;

size_t sort_workspace_bytes(uint64_t n);
int sort_num_passes(int key_bits);
cudaError_t sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                       uint64_t n, int key_bits, bool iota_vals, void* workspace, int num_sms,
                       cudaStream_t stream, bool* result_in_b);

}  // namespace nprsph
