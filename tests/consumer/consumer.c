/* A plain C consumer of include/nprsph.h: what a maintainer's headless driver would do with the
 * library in place of Main.cpp's SPH block (init_particles :510-539, display() :291-305, keyboard
 * 'p' :466-469).  Prints an FNV-1a checksum of the downloaded records; tests/test_gpu_consumer.py
 * compares it with the same run through the ctypes harness. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "nprsph.h"

static uint64_t fnv1a(const void* data, size_t bytes) {
    const unsigned char* p = (const unsigned char*)data;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < bytes; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 5;
    if (nprsph_abi_version() != NPRSPH_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 2; }
    nprsph_config cfg;
    nprsph_config_default(&cfg);
    nprsph_ctx* ctx = NULL;
    int rc = nprsph_create(&cfg, &ctx);
    if (rc != NPRSPH_OK) { fprintf(stderr, "nprsph_create: %d %s\n", rc, nprsph_last_error(NULL)); return 1; }
    /* the context starts with the reference's 10 x 100 x 10 block, paused (Main.cpp:87) */
    const uint64_t n = nprsph_num_particles(ctx);
    nprsph_particle* rec = (nprsph_particle*)malloc(n * sizeof *rec);
    float* pos = (float*)malloc(n * 4 * sizeof(float));
    if (!rec || !pos) return 1;
    rc = nprsph_step(ctx, 1);                        /* no-op while paused */
    if (rc == NPRSPH_OK) rc = nprsph_toggle_pause(ctx);            /* 'p' */
    if (rc == NPRSPH_OK) rc = nprsph_step(ctx, steps);
    if (rc == NPRSPH_OK) rc = nprsph_download_particles(ctx, rec, n);
    if (rc == NPRSPH_OK) rc = nprsph_download_positions(ctx, pos, n, 0);
    if (rc != NPRSPH_OK) { fprintf(stderr, "error %d: %s\n", rc, nprsph_last_error(ctx)); return 1; }
    int same = 1;
    for (uint64_t i = 0; i < n && same; i++)
        for (int k = 0; k < 4; k++) same = same && (pos[4 * i + k] == rec[i].pos[k] || (pos[4 * i + k] != pos[4 * i + k] && rec[i].pos[k] != rec[i].pos[k]));
    nprsph_stats st;
    nprsph_get_stats(ctx, &st);
    printf("particles %llu steps %llu cell_subdiv %d positions_match %d checksum %016llx\n", (unsigned long long)n,
           (unsigned long long)st.steps_done, st.cell_subdiv, same, (unsigned long long)fnv1a(rec, n * sizeof *rec));
    free(rec); free(pos);
    return nprsph_destroy(ctx) == NPRSPH_OK && same ? 0 : 1;
}
