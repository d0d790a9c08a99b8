// Yard-stick for the hand-written onesweep sort: CUB's DeviceRadixSort on the same
// problem -- 2^24 (key, slot) pairs, 27 significant key bits -- timed with CUDA events.  NOT linked
// into the product; build and run by hand:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/cub_sort_yardstick.cu -o /tmp/cub_sort && /tmp/cub_sort
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdint>
#include <vector>

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : (size_t)1 << 24;
    const int bits = argc > 2 ? atoi(argv[2]) : 27;
    uint32_t *k0, *k1, *v0, *v1;
    cudaMalloc(&k0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
    std::vector<uint32_t> h(n);
    for (int mode = 0; mode < 2; mode++) {
        // mode 0: uniformly random keys; mode 1: almost sorted (what a step sees: <1 % of the particles change cell)
        uint64_t s = 88172645463325252ull;
        for (size_t i = 0; i < n; i++) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            h[i] = mode == 0 ? (uint32_t)(s & ((1u << bits) - 1u))
                             : (uint32_t)((i * 6) & ((1u << bits) - 1u)) + ((s & 127) == 0 ? 3u : 0u);
        }
        cudaMemcpy(k0, h.data(), n * 4, cudaMemcpyHostToDevice);
        cub::DoubleBuffer<uint32_t> dk(k0, k1), dv(v0, v1);
        size_t ws_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, ws_bytes, dk, dv, (int)n, 0, bits);
        void* ws; cudaMalloc(&ws, ws_bytes);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f, sum = 0.f;
        const int reps = 20;
        for (int r = 0; r < reps + 3; r++) {
            cudaMemcpy(k0, h.data(), n * 4, cudaMemcpyHostToDevice);
            cub::DoubleBuffer<uint32_t> a(k0, k1), b(v0, v1);
            cudaEventRecord(e0);
            cub::DeviceRadixSort::SortPairs(ws, ws_bytes, a, b, (int)n, 0, bits);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 3) { sum += ms; if (ms < best) best = ms; }
        }
        const int passes = (bits + 7) / 8;
        printf("{\"cub_sort_pairs\": \"%s\", \"n\": %zu, \"key_bits\": %d, \"mean_ms\": %.4f, \"best_ms\": %.4f, "
               "\"digit_passes_8bit\": %d, \"bytes_per_pair\": %d, \"gbps_of_4+16P_bytes\": %.1f}\n",
               mode == 0 ? "random keys" : "almost sorted keys", n, bits, sum / reps, best, passes, 4 + 16 * passes,
               (double)n * (4 + 16 * passes) / (sum / reps * 1e-3) / 1e9);
        cudaFree(ws);
    }
    return 0;
}
