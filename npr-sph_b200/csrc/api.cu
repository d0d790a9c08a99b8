// api.cu -- host runtime and C ABI of libnprsph.so (include/nprsph.h).
//
// Replaces the SPH-driver part of the reference's Main.cpp: particle/UBO creation
// (Main.cpp:510-539,631-641), the per-frame dispatch block (Main.cpp:291-305), parameter
// upload (Main.cpp:274-278) and the p/r keys (Main.cpp:454-476).  Everything runs on one CUDA
// stream; there is no CPU implementation of any pass behind this API.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "context.cuh"

using namespace nprsph;

// CUDA-GL interop entry points of the CUDA runtime, declared by hand because this image has
// no <GL/gl.h> for cuda_gl_interop.h to include.
extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource** resource,
                                                    unsigned int buffer, unsigned int flags);

namespace {
thread_local std::string g_create_error;
}

namespace nprsph {

int fail(nprsph_ctx* c, int code, const char* fmt, const char* detail) {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, detail);
    if (c) { c->err = buf; if (code == NPRSPH_ERR_CUDA) c->sticky = code; }
    else g_create_error = buf;
    return code;
}

}  // namespace nprsph

namespace {

// Smallest fp32 t with sqrtf(t) >= h  (so `length(d) < h` == `r2 < t`; sqrtf is monotone).
float r2_threshold(float h) {
    if (!(h > 0.0f)) return 0.0f;
    if (isinf(h)) return INFINITY;
    float t = h * h;
    while (t > 0.0f && sqrtf(nextafterf(t, 0.0f)) >= h) t = nextafterf(t, 0.0f);
    while (sqrtf(t) < h) t = nextafterf(t, INFINITY);
    return t;
}

// Uniform-grid definition (DESIGN.md "Grid"): cell = h*(1+2^-14)/subdiv, enlarged when an axis
// would exceed 16384 cells or the table would exceed max_cells.
int setup_grid(const nprsph_ctx* c, float h, int slab_ranks, GridDev* g, float* cell_size) {
    int k = effective_subdiv(c);
    if (!(h > 0.0f) || isinf(h) || k < 1 || k > 4) return -1;
    double ext[3];
    for (int a = 0; a < 3; a++) {
        double e = (double)c->bounds.upper[a] - (double)c->bounds.lower[a];
        if (!(e == e) || isinf(e)) return -2;
        ext[a] = e > 0.0 ? e : 0.0;
    }
    // slab mode: the cap applies to a rank's share of the table, not to the global grid
    double max_cells = (double)(c->cfg.max_cells ? c->cfg.max_cells : (1u << 28)) * (double)slab_ranks;
    // cell = base * (1 + widen), widen = 2^-14.  Two particles closer than h (fp32 predicate, rounding
    // <= 2^-22 relative) are fewer than `reach` cells apart; the cell coordinate itself is computed in
    // fp64 (common.cuh:cell_coord, error ~ dim * 2^-52 cells), so the cell size does not have to grow
    // with the grid.  (An fp32 coordinate needed widen >= dim_max * 2^-21: the cell size then depended
    // on the box, i.e. on the number of GPUs, and a lattice block slipped against the cells inside it.)
    double base = (double)h / (double)k, widen = 1.0 / 16384.0, cell = 0.0;
    double dims[3] = {1, 1, 1};
    for (int iter = 0; iter < 64; iter++) {
        cell = base * (1.0 + widen);
        double dmax = 1.0;
        for (int a = 0; a < 3; a++) {
            dims[a] = floor(ext[a] / cell) + 1.0;
            if (dims[a] > dmax) dmax = dims[a];
        }
        if (dmax > 16384.0) { base *= dmax / 16383.0 * 1.0001; continue; }
        const double total = dims[0] * dims[1] * dims[2];
        if (total > max_cells) { base *= cbrt(total / max_cells) * 1.0001; continue; }
        break;
    }
    for (int a = 0; a < 3; a++) { g->lo[a] = c->bounds.lower[a]; g->dim[a] = (int)dims[a]; }
    g->inv_cell_d = 1.0 / cell;
    g->inv_cell = (float)g->inv_cell_d;      // fp32 copy: culling coordinates of the walks only
    g->pad0 = 0;
    g->reach = k;
    g->num_cells = (uint32_t)((int64_t)g->dim[0] * g->dim[1] * g->dim[2]);   // (wraps in slab mode: unused there)
    g->x_off = 0;
    g->dimx_global = g->dim[0];
    *cell_size = (float)cell;
    return 0;
}

}  // namespace

// cell_subdiv 0 = automatic: the reference's block sits on a lattice of spacing PARTICLE_RADIUS
// (make_grid, Main.cpp:488-505) and h = smoothing_coeff * PARTICLE_RADIUS (rho_pres_comp.glsl:40), so
// cell = h / round(smoothing_coeff) holds about one particle: 4 at the reference's default, 2 for the
// h = 2s dam break.  Columns of a walk then stay within the 16 hit bits of a pair record.
int nprsph::effective_subdiv(const nprsph_ctx* c) {
    if (c->cfg.cell_subdiv > 0) return c->cfg.cell_subdiv;
    const float k = roundf(c->consts.smoothing_coeff);
    return k >= 4.0f ? 4 : k >= 1.0f ? (int)k : 1;
}

int nprsph::ensure_records(nprsph_ctx* c, uint64_t slots, int reach) {
    size_t need = 0;
    // A column record holds 16 hit bits per target.  With cells much larger than the lattice spacing
    // (the GUI's smoothing slider, 7..10 radii, at the finest subdivision 4: 5-15 particles per cell)
    // every column of every walk overflows: the density pass would write records nobody can use and the
    // whole force pass would run as re-tests out of the deferred queue.  Then there are no records at
    // all and both passes take their plain paths (k_rho without records, k_force_scan with a full grid).
    const bool cells_too_coarse = c->consts.smoothing_coeff > 1.5f * (float)effective_subdiv(c);
    if (!(c->cfg.flags & NPRSPH_FLAG_NO_HITMASK) && slots && records_fit(reach, slots) && !cells_too_coarse)
        need = rec_buffer_words(slots, rec_cols_of(reach));
    if (need == c->hitmask_words) return NPRSPH_OK;
    CK(c, cudaStreamSynchronize(c->stream));
    c->hitmask_words = 0;
    CK(c, realloc_dev(c->hitmask, need));
    c->hitmask_words = need;
    c->mask_valid = false;
    return NPRSPH_OK;
}

int nprsph::refresh_params(nprsph_ctx* c) {
    if (!c->params_dirty) return NPRSPH_OK;
    const nprsph_config& cf = c->cfg;
    SphDev s;
    s.h = c->consts.smoothing_coeff * cf.particle_radius;       // rho_pres_comp.glsl:40
    s.h2 = s.h * s.h;
    s.r2_max = r2_threshold(s.h);
    s.one = 1.0f;
    s.zero = 0.0f;
    const double h = (double)s.h, pi = (double)cf.pi, m = (double)c->consts.mass;
    s.rho_coef = (float)(m * 315.0 / (64.0 * pi * pow(h, 9.0)));            // :52
    s.pres_coef = (float)(m * 45.0 / (2.0 * pi * pow(h, 6.0)));             // force_comp.glsl:41,59
    s.visc_coef = (float)((double)c->consts.visc * m * 45.0 / (pi * pow(h, 6.0)));   // :42,60,63
    s.gas_const = cf.gas_const;
    s.rest_rho = c->consts.resting_rho;
    for (int a = 0; a < 3; a++) {
        s.g[a] = cf.gravity[a];
        s.lower[a] = c->bounds.lower[a];
        s.upper[a] = c->bounds.upper[a];
    }
    s.damping = cf.damping;
    s.dt = cf.dt;

    GridDev g;
    float cell_size;
    if (setup_grid(c, s.h, slab_world(c), &g, &cell_size))
        return fail(c, NPRSPH_ERR_INVALID, "cannot build a grid: smoothing length or bounds invalid%s");
    {   // column cull threshold: h in cell units plus a margin above twice the fp32 error of the
        // cell coordinates (<= dim * 2^-23 cells each)
        int dmax = g.dim[0] > g.dim[1] ? g.dim[0] : g.dim[1];
        if (g.dim[2] > dmax) dmax = g.dim[2];
        const float margin = fmaxf(1.0f / 256.0f, (float)dmax / 1048576.0f);
        const float hc = s.h * g.inv_cell + margin;
        s.cull2 = hc * hc;
    }
    c->mask_valid = false;
    if (memcmp(&g, &c->grid, sizeof g) != 0) {
        c->keys_valid = false; c->grid_valid = false;
        CK(c, cudaMemsetAsync(c->gap_count + 4, 0xFF, 8, c->stream));   // cell table contents are void
    }
    const size_t need = c->dist ? 0 : (size_t)g.num_cells + 2;   // slab mode sizes its own local table
    if (need > c->cell_cap) {
        if (c->cell_start) CK(c, cudaFree(c->cell_start));
        c->cell_start = nullptr; c->cell_cap = 0;
        CK(c, cudaMalloc(&c->cell_start, need * sizeof(uint32_t)));
        c->cell_cap = need;
    }
    const size_t gaps = c->dist ? 0 : gap_list_capacity(g.num_cells, c->cap ? c->cap : 1);
    if (gaps > c->gap_cap) {
        if (c->gap_list) CK(c, cudaFree(c->gap_list));
        c->gap_list = nullptr; c->gap_cap = 0;
        CK(c, cudaMalloc(&c->gap_list, gaps * sizeof(uint4)));
        c->gap_cap = gaps;
    }
    if (!c->dist) {              // slab mode sizes its records in alloc_slab
        int rc = ensure_records(c, c->cap, g.reach);
        if (rc) return rc;
    }
    c->grid = g;
    c->sph = s;
    c->cell_size = cell_size;
    int bits = 1;
    while (bits < 32 && (1ull << bits) <= (uint64_t)g.num_cells) bits++;    // keys 0..num_cells
    c->key_bits = bits;
    c->params_dirty = false;
    return NPRSPH_OK;
}

namespace {

int ensure_capacity(nprsph_ctx* c, uint64_t n) {
    if (n >= (1ull << 30)) return fail(c, NPRSPH_ERR_INVALID, "at most 2^30-1 particles per context%s");
    if (n <= c->cap) return NPRSPH_OK;
    CK(c, cudaStreamSynchronize(c->stream));
    c->cap = 0;
    { float4* a = (float4*)c->aos; CK(c, realloc_dev(a, n * 4)); c->aos = a; }
    for (int b = 0; b < 2; b++) {
        CK(c, realloc_dev(c->pos[b], n));
        CK(c, realloc_dev(c->vel[b], n));
        CK(c, realloc_dev(c->frc[b], n));
        CK(c, realloc_dev(c->keys[b], n));
        CK(c, realloc_dev(c->vals[b], n));
    }
    if (c->cfg.flags & NPRSPH_FLAG_COUNT_NEIGHBOURS) {
        CK(c, realloc_dev(c->counts_rho, n));
        CK(c, realloc_dev(c->counts_force, n));
    }
    { char* w = (char*)c->sort_ws; CK(c, realloc_dev(w, sort_workspace_bytes(n))); c->sort_ws = w; }
    c->hitmask_words = 0;       // sized with the grid (refresh_params -> ensure_records)
    CK(c, realloc_dev(c->hitmask, (size_t)0));
    c->cap = n;
    c->params_dirty = true;     // gap-list capacity depends on cap
    return NPRSPH_OK;
}

// Bring the cell-ordered state and the cell table up to date with the current positions.
int ensure_grid(nprsph_ctx* c, bool with_force, cudaEvent_t* ev /* 4 events or null */) {
    int rc = refresh_params(c);
    if (rc) return rc;
    if (c->grid_valid || c->n == 0) {
        if (ev) for (int i = 0; i < 4; i++) CK(c, cudaEventRecord(ev[i], c->stream));
        return NPRSPH_OK;
    }
    const uint32_t n = (uint32_t)c->n;
    if (ev) CK(c, cudaEventRecord(ev[0], c->stream));
    if (!c->keys_valid) {
        launch_keys(c->pos[c->cur], c->keys[0], n, c->grid, c->stream);
        c->keys_valid = true;
    }
    if (ev) CK(c, cudaEventRecord(ev[1], c->stream));
    bool in_b = false;
    CK(c, sort_pairs(c->keys[0], c->vals[0], c->keys[1], c->vals[1], n, c->key_bits, true,
                     c->sort_ws, c->num_sms, c->stream, &in_b));
    c->sorted_keys = in_b ? c->keys[1] : c->keys[0];
    c->last_perm = in_b ? c->vals[1] : c->vals[0];
    if (ev) CK(c, cudaEventRecord(ev[2], c->stream));
    const int nxt = 1 - c->cur;
    launch_reorder_cells(c->sorted_keys, c->last_perm, c->pos[c->cur], c->vel[c->cur],
                         c->frc[c->cur], c->pos[nxt], c->vel[nxt], c->frc[nxt], c->cell_start,
                         c->grid.num_cells, n, c->gap_list, c->gap_count, with_force, c->num_sms,
                         c->stream);
    if (ev) CK(c, cudaEventRecord(ev[3], c->stream));
    c->cur = nxt;
    c->grid_valid = true;
    c->mask_valid = false;
    CK(c, cudaGetLastError());
    return NPRSPH_OK;
}

int run_rho(nprsph_ctx* c, bool write_pressure) {
    launch_rho(c->pos[c->cur], c->vel[c->cur], write_pressure ? c->frc[c->cur] : nullptr,
               c->cell_start, 0u, (uint32_t)c->n, c->grid, c->sph, c->counts_rho, c->hitmask,
               (uint32_t)c->cap, c->stream);
    c->mask_valid = c->hitmask != nullptr;
    c->aos_stale = true;
    return NPRSPH_OK;
}

int run_force(nprsph_ctx* c) {
    launch_force(c->pos[c->cur], c->vel[c->cur], c->frc[c->cur], c->cell_start, 0u, (uint32_t)c->n,
                 c->grid, c->sph, c->counts_force, c->mask_valid ? c->hitmask : nullptr,
                 (uint32_t)c->cap, c->stream);
    c->aos_stale = true;
    return NPRSPH_OK;
}

int run_integrate(nprsph_ctx* c) {
    // keys for the next step are written where the sort expects its input
    launch_integrate(c->pos[c->cur], c->vel[c->cur], c->frc[c->cur], c->keys[0], (uint32_t)c->n,
                     c->grid, c->sph, c->colliders, c->stream);
    c->keys_valid = true;
    c->grid_valid = false;
    c->mask_valid = false;
    c->sorted_keys = nullptr;
    c->aos_stale = true;
    return NPRSPH_OK;
}

// Passes 2 + 3 of a full step.  With the density pass's column records at hand both run in ONE
// launch: the force kernel integrates its own two particles from registers and writes force,
// position and velocity into the other buffer set (neighbours still gather the old positions), so
// the stand-alone k_integrate launch and its 96 B/particle of traffic disappear.  fused_out
// (nullable) reports which form ran.
int run_force_integrate(nprsph_ctx* c, bool* fused_out, cudaEvent_t between /* nullable */) {
    const int nxt = 1 - c->cur;
    const bool fused = !(c->cfg.flags & NPRSPH_FLAG_NO_FUSE) &&
        launch_force_integrate(c->pos[c->cur], c->vel[c->cur], c->frc[nxt], c->cell_start, (uint32_t)c->n,
                               c->grid, c->sph, c->counts_force, c->mask_valid ? c->hitmask : nullptr,
                               (uint32_t)c->cap, c->pos[nxt], c->vel[nxt], c->keys[0], c->colliders, c->stream);
    if (fused_out) *fused_out = fused;
    if (!fused) {
        run_force(c);
        if (between) CK(c, cudaEventRecord(between, c->stream));
        return run_integrate(c);
    }
    if (between) CK(c, cudaEventRecord(between, c->stream));
    c->cur = nxt;
    c->keys_valid = true;
    c->grid_valid = false;
    c->mask_valid = false;
    c->sorted_keys = nullptr;
    c->aos_stale = true;
    return NPRSPH_OK;
}

int publish(nprsph_ctx* c) {
    if (!c->aos_stale || c->n == 0) return NPRSPH_OK;
    launch_publish(c->pos[c->cur], c->vel[c->cur], c->frc[c->cur], c->aos, (uint32_t)c->n, c->stream);
    CK(c, cudaGetLastError());
    c->aos_stale = false;
    return NPRSPH_OK;
}

// after the AoS buffer was (re)written by upload or a scene kernel
int adopt_aos(nprsph_ctx* c) {
    c->cur = 0;
    c->gl_seeded = false;          // every lane of the records may have changed
    launch_import(c->aos, c->pos[0], c->vel[0], c->frc[0], (uint32_t)c->n, c->stream);
    CK(c, cudaGetLastError());
    c->keys_valid = false;
    c->grid_valid = false;
    c->mask_valid = false;
    c->aos_stale = false;
    c->sorted_keys = nullptr;
    return NPRSPH_OK;
}

__device__ __forceinline__ uint32_t hash32(uint32_t seed, uint32_t idx) {
    uint32_t x = seed ^ (idx * 0x9E3779B9u);
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

// make_grid() + init_particles(), Main.cpp:488-521, generated on the device.
__global__ void __launch_bounds__(256)
k_scene_block(float4* __restrict__ aos, int nx, int ny, int nz, float spacing, float ox, float oy,
              float oz, float jitter, uint32_t seed) {
    const uint64_t n = (uint64_t)nx * ny * nz;
    const uint64_t idx = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    const int k = (int)(idx % nz);
    const int j = (int)((idx / nz) % ny);
    const int i = (int)(idx / ((uint64_t)nz * ny));
    float x = __fadd_rn(__fmul_rn((float)i, spacing), ox);
    float y = __fadd_rn(__fmul_rn((float)j, spacing), oy);
    float z = __fadd_rn(__fmul_rn((float)k, spacing), oz);
    if (jitter > 0.0f) {
        float c[3] = {x, y, z};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const uint32_t u = hash32(seed, (uint32_t)(3 * idx + a));
            const float f = __fmul_rn((float)(u >> 8), 1.0f / 16777216.0f);
            c[a] = __fadd_rn(c[a], __fmul_rn(__fsub_rn(__fmul_rn(2.0f, f), 1.0f), jitter));
        }
        x = c[0]; y = c[1]; z = c[2];
    }
    aos[4 * idx + 0] = make_float4(x, y, z, 1.0f);
    aos[4 * idx + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    aos[4 * idx + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
    aos[4 * idx + 3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

int build_scene(nprsph_ctx* c) {
    const uint64_t n = (uint64_t)c->scene.nx * c->scene.ny * c->scene.nz;
    int rc = ensure_capacity(c, n);
    if (rc) return rc;
    c->n = n;
    if (n) {
        k_scene_block<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
            (float4*)c->aos, c->scene.nx, c->scene.ny, c->scene.nz, c->scene.spacing,
            c->scene.origin[0], c->scene.origin[1], c->scene.origin[2], c->scene.jitter, c->scene.seed);
        CK(c, cudaGetLastError());
    }
    return adopt_aos(c);
}

bool config_ok(const nprsph_config* cfg) {
    return cfg && cfg->struct_size == sizeof(nprsph_config) && cfg->cell_subdiv >= 0 && cfg->cell_subdiv <= 4;
}

}  // namespace

// ================================ C ABI ==========================================================
extern "C" {

int nprsph_abi_version(void) { return NPRSPH_ABI_VERSION; }

void nprsph_config_default(nprsph_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = sizeof *cfg;
    cfg->device = 0;
    cfg->stream = nullptr;
    cfg->particle_radius = 0.005f;
    cfg->gas_const = 2000.0f;
    cfg->gravity[0] = 0.0f; cfg->gravity[1] = -9806.65f; cfg->gravity[2] = 0.0f;
    cfg->damping = 0.3f;
    cfg->dt = 1.0f / 10000.0f;
    cfg->pi = 3.141592741f;
    cfg->cell_subdiv = 0;       // automatic: effective_subdiv()
    cfg->max_cells = 0;
    cfg->flags = 0;
}

const char* nprsph_last_error(const nprsph_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int nprsph_create(const nprsph_config* cfg, nprsph_ctx** out) {
    if (!out) return NPRSPH_ERR_INVALID;
    *out = nullptr;
    if (!config_ok(cfg)) return fail(nullptr, NPRSPH_ERR_INVALID, "bad nprsph_config (struct_size / cell_subdiv)%s");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, NPRSPH_ERR_CUDA, "no CUDA device: %s (libnprsph has no CPU path)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (cfg->device < 0 || cfg->device >= count) return fail(nullptr, NPRSPH_ERR_INVALID, "bad device ordinal%s");
    CK(nullptr, cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(nullptr, NPRSPH_ERR_UNSUPPORTED, "libnprsph is built for sm_100a only; device is %s", prop.name);
    nprsph_ctx* c = new (std::nothrow) nprsph_ctx();
    if (!c) return fail(nullptr, NPRSPH_ERR_NOMEM, "out of host memory%s");
    c->cfg = *cfg;
    {
        const char* off = getenv("NPRSPH_NO_GRAPH");
        c->graph_off = (cfg->flags & NPRSPH_FLAG_NO_GRAPH) || (off && off[0] && off[0] != '0');
    }
    c->num_sms = prop.multiProcessorCount;
    // ConstantsData / BoundaryData defaults, Main.cpp:110-122
    c->consts = {0.02f, 4.0f, 3000.0f, 1000.0f};
    c->bounds = {{0.5f, 1.0f, 0.5f, 1.0f}, {-0.1f, -0.35f, -0.1f, 1.0f}};
    // make_grid(): 10 x 100 x 10 at PARTICLE_RADIUS spacing, Main.cpp:488-505
    c->scene = {10, 100, 10, cfg->particle_radius, {0.f, 0.f, 0.f}, 0.f, 0u};
    memset(&c->grid, 0, sizeof c->grid);
    memset(&c->sph, 0, sizeof c->sph);
    if (cfg->stream) { c->stream = (cudaStream_t)cfg->stream; c->own_stream = false; }
    else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail(nullptr, NPRSPH_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        c->own_stream = true;
    }
    // [0] medium-gap count, [1] huge-gap count, [2..3] NaN counter, [4..5] tail state of the cell table,
    // [8..] huge-gap list (common.cuh:push_gap)
    e = cudaMalloc(&c->gap_count, GAP_CTL_BYTES);
    if (e == cudaSuccess) e = cudaMemset(c->gap_count, 0xFF, GAP_CTL_BYTES);
    if (e != cudaSuccess) { nprsph_destroy(c); return fail(nullptr, NPRSPH_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
    int rc = build_scene(c);      // initOpenGL() -> init_particles(), Main.cpp:567
    if (rc) { g_create_error = c->err; nprsph_destroy(c); return rc; }
    *out = c;
    return NPRSPH_OK;
}

int nprsph_destroy(nprsph_ctx* c) {
    if (!c) return NPRSPH_OK;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->gl_res) cudaGraphicsUnregisterResource(c->gl_res);
    if (c->step_graph) cudaGraphExecDestroy(c->step_graph);
    if (c->dist) dist_destroy(c);
    if (c->h2d_stream) { cudaStreamSynchronize(c->h2d_stream); cudaStreamDestroy(c->h2d_stream); }
    if (c->d2h_stream) { cudaStreamSynchronize(c->d2h_stream); cudaStreamDestroy(c->d2h_stream); }
    for (int i = 0; i < nprsph_ctx::STAGE_CHUNKS; i++) if (c->ev_chunk[i]) cudaEventDestroy(c->ev_chunk[i]);
    for (int i = 0; i < 2; i++) { if (c->ev_imported[i]) cudaEventDestroy(c->ev_imported[i]); cudaFree(c->stage_in[i]); }
    if (c->ev_published) cudaEventDestroy(c->ev_published);
    if (c->ev_copied) cudaEventDestroy(c->ev_copied);
    cudaFree(c->stage_pos);
    cudaFree(c->aos);
    for (int b = 0; b < 2; b++) {
        cudaFree(c->pos[b]); cudaFree(c->vel[b]); cudaFree(c->frc[b]);
        cudaFree(c->keys[b]); cudaFree(c->vals[b]);
    }
    cudaFree(c->counts_rho); cudaFree(c->counts_force); cudaFree(c->sort_ws); cudaFree(c->hitmask);
    cudaFree(c->cell_start); cudaFree(c->gap_list); cudaFree(c->gap_count);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return NPRSPH_OK;
}

// ---- parameters ------------------------------------------------------------------------------------
int nprsph_set_constants(nprsph_ctx* c, const nprsph_constants* k) {
    if (!c || !k) return NPRSPH_ERR_INVALID;
    c->consts = *k;
    c->params_dirty = true;
    return NPRSPH_OK;
}
int nprsph_get_constants(const nprsph_ctx* c, nprsph_constants* k) {
    if (!c || !k) return NPRSPH_ERR_INVALID;
    *k = c->consts;
    return NPRSPH_OK;
}
int nprsph_set_boundary(nprsph_ctx* c, const nprsph_boundary* b) {
    if (!c || !b) return NPRSPH_ERR_INVALID;
    c->bounds = *b;
    c->params_dirty = true;
    return NPRSPH_OK;
}
int nprsph_get_boundary(const nprsph_ctx* c, nprsph_boundary* b) {
    if (!c || !b) return NPRSPH_ERR_INVALID;
    *b = c->bounds;
    return NPRSPH_OK;
}
// Static colliders: kernel parameters of the integrate kernels, so they take effect at the next
// step like the uniform blocks (sendUniforms, Main.cpp:274-278).
int nprsph_set_colliders(nprsph_ctx* c, const nprsph_collider* list, int n) {
    if (!c || n < 0 || (n && !list)) return NPRSPH_ERR_INVALID;
    if (n > NPRSPH_MAX_COLLIDERS) return fail(c, NPRSPH_ERR_INVALID, "more than NPRSPH_MAX_COLLIDERS colliders%s");
    ColliderSet cs = {};
    for (int i = 0; i < n; i++) {
        const nprsph_collider& k = list[i];
        if (k.kind == NPRSPH_COLLIDER_SPHERE) {
            if (!(k.b[0] > 0.0f)) return fail(c, NPRSPH_ERR_INVALID, "sphere collider needs a positive radius%s");
        } else if (k.kind == NPRSPH_COLLIDER_BOX) {
            for (int a = 0; a < 3; a++)
                if (!(k.a[a] < k.b[a])) return fail(c, NPRSPH_ERR_INVALID, "box collider needs lower < upper%s");
        } else {
            return fail(c, NPRSPH_ERR_INVALID, "unknown collider kind%s");
        }
        cs.a[i] = make_float4(k.a[0], k.a[1], k.a[2], (float)k.kind);
        cs.b[i] = make_float4(k.b[0], k.b[1], k.b[2], 0.0f);
    }
    cs.n = n;
    c->colliders = cs;
    return NPRSPH_OK;
}
int nprsph_get_colliders(const nprsph_ctx* c, nprsph_collider* out, int cap) {
    if (!c || cap < 0 || (cap && !out)) return NPRSPH_ERR_INVALID;
    const int n = c->colliders.n < cap ? c->colliders.n : cap;
    for (int i = 0; i < n; i++) {
        const float4 a = c->colliders.a[i], b = c->colliders.b[i];
        out[i].kind = (uint32_t)a.w;
        out[i].a[0] = a.x; out[i].a[1] = a.y; out[i].a[2] = a.z;
        out[i].b[0] = b.x; out[i].b[1] = b.y; out[i].b[2] = b.z;
        out[i].reserved = 0.0f;
    }
    return c->colliders.n;
}

// The "Constants Window" of the reference GUI (Main.cpp:240-247): four ImGui::SliderFloat widgets
// bound to ConstantsData.  A slider edit clamps to the widget's range; the stored defaults
// (Main.cpp:110-116) may lie outside it (smoothing_coeff 4 vs a 7..10 slider) until touched.
static const nprsph_slider g_sliders[NPRSPH_NUM_SLIDERS] = {
    {"Mass", 0.01f, 0.1f, 0.02f},                   // Main.cpp:242
    {"Smoothing", 7.0f, 10.0f, 4.0f},               // Main.cpp:243
    {"Viscosity", 1000.0f, 5000.0f, 3000.0f},       // Main.cpp:244
    {"Resting Density", 1000.0f, 5000.0f, 1000.0f}, // Main.cpp:245
};
int nprsph_slider_info(int id, nprsph_slider* out) {
    if (id < 0 || id >= NPRSPH_NUM_SLIDERS || !out) return NPRSPH_ERR_INVALID;
    *out = g_sliders[id];
    return NPRSPH_OK;
}
int nprsph_set_slider(nprsph_ctx* c, int id, float value) {
    if (!c) return NPRSPH_ERR_INVALID;
    if (id < 0 || id >= NPRSPH_NUM_SLIDERS || value != value) return fail(c, NPRSPH_ERR_INVALID, "bad slider id or value%s");
    const nprsph_slider& s = g_sliders[id];
    const float v = value < s.min ? s.min : (value > s.max ? s.max : value);   // ImGui clamps the edit
    float* field[NPRSPH_NUM_SLIDERS] = {&c->consts.mass, &c->consts.smoothing_coeff, &c->consts.visc,
                                        &c->consts.resting_rho};
    *field[id] = v;
    c->params_dirty = true;
    return NPRSPH_OK;
}

int nprsph_set_config(nprsph_ctx* c, const nprsph_config* cfg) {
    if (!c) return NPRSPH_ERR_INVALID;
    if (!config_ok(cfg)) return fail(c, NPRSPH_ERR_INVALID, "bad nprsph_config%s");
    if (cfg->device != c->cfg.device || cfg->stream != c->cfg.stream)
        return fail(c, NPRSPH_ERR_INVALID, "device and stream are fixed at create time%s");
    if (cfg->flags != c->cfg.flags)
        return fail(c, NPRSPH_ERR_INVALID, "flags are fixed at create time%s");
    c->cfg = *cfg;
    c->params_dirty = true;
    return NPRSPH_OK;
}
int nprsph_get_config(const nprsph_ctx* c, nprsph_config* cfg) {
    if (!c || !cfg) return NPRSPH_ERR_INVALID;
    *cfg = c->cfg;
    return NPRSPH_OK;
}

// ---- particle buffer -------------------------------------------------------------------------------
int nprsph_scene_block(nprsph_ctx* c, int nx, int ny, int nz, float spacing, const float origin[3],
                       float jitter, uint32_t seed) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (nx < 0 || ny < 0 || nz < 0) return fail(c, NPRSPH_ERR_INVALID, "negative block size%s");
    c->scene.nx = nx; c->scene.ny = ny; c->scene.nz = nz;
    c->scene.spacing = spacing;
    for (int a = 0; a < 3; a++) c->scene.origin[a] = origin ? origin[a] : 0.0f;
    c->scene.jitter = jitter;
    c->scene.seed = seed;
    return build_scene(c);
}

int nprsph_upload_particles(nprsph_ctx* c, const nprsph_particle* host, uint64_t n) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (n && !host) return fail(c, NPRSPH_ERR_INVALID, "null host buffer%s");
    int rc = ensure_capacity(c, n);
    if (rc) return rc;
    c->n = n;
    if (n) CK(c, cudaMemcpyAsync(c->aos, host, n * sizeof(nprsph_particle), cudaMemcpyHostToDevice, c->stream));
    return adopt_aos(c);
}

int nprsph_download_particles(nprsph_ctx* c, nprsph_particle* host, uint64_t n) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (n > c->n) return fail(c, NPRSPH_ERR_INVALID, "download larger than the particle buffer%s");
    if (n && !host) return fail(c, NPRSPH_ERR_INVALID, "null host buffer%s");
    int rc = publish(c);
    if (rc) return rc;
    if (n) CK(c, cudaMemcpyAsync(host, c->aos, n * sizeof(nprsph_particle), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return NPRSPH_OK;
}

int nprsph_device_particles(nprsph_ctx* c, void** device_ptr, uint64_t* n) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (!device_ptr) return NPRSPH_ERR_INVALID;
    int rc = publish(c);
    if (rc) return rc;
    *device_ptr = c->aos;
    if (n) *n = c->n;
    return NPRSPH_OK;
}

// ---- streaming interface ---------------------------------------------------------------------------
// The reference uploads the particle buffer once and never reads it back; a host application that
// feeds state in (restart, emitters, a coupled solver) and draws the result needs two things per
// frame: the INPUTS of a step (positions and velocities -- force, density and pressure are outputs of
// the passes) and the positions the renderer consumes (attribute 0, Main.cpp:533-535).
static int ensure_staging(nprsph_ctx* c) {
    if (!c->h2d_stream) {
        CK(c, cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
        CK(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < nprsph_ctx::STAGE_CHUNKS; i++) CK(c, cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) CK(c, cudaEventCreateWithFlags(&c->ev_imported[i], cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_published, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
    }
    if (c->stage_cap < c->n) {
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, cudaStreamSynchronize(c->h2d_stream));
        CK(c, cudaStreamSynchronize(c->d2h_stream));
        c->stage_cap = 0;
        for (int i = 0; i < 2; i++) CK(c, realloc_dev(c->stage_in[i], 2 * c->n));
        CK(c, realloc_dev(c->stage_pos, c->n));
        c->stage_cap = c->n;
        // fresh events: nothing recorded yet counts as complete
    }
    return NPRSPH_OK;
}

int nprsph_upload_state(nprsph_ctx* c, const float* pos4, const float* vel4, uint64_t n) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (n != c->n) return fail(c, NPRSPH_ERR_INVALID, "upload_state replaces the state of the existing particles: n must equal "
                               "nprsph_num_particles() (nprsph_upload_particles changes the count)%s");
    if (!n) return NPRSPH_OK;
    if (!pos4 || !vel4) return fail(c, NPRSPH_ERR_INVALID, "null host buffer%s");
    int rc = ensure_staging(c);
    if (rc) return rc;
    // Chunked: the copy stream brings chunk k+1 in while the context's stream converts chunk k into
    // the working arrays; the staging buffers alternate between calls, so this call's copies may run
    // while the previous step still computes.
    const int b = c->stage_cur;
    c->stage_cur ^= 1;
    float4* sp = c->stage_in[b];
    float4* sv = sp + n;
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
        launch_import_state(sp + first, sv + first, (uint32_t)first, (uint32_t)cnt, c->pos[0], c->vel[0], c->frc[0], c->stream);
    }
    CK(c, cudaEventRecord(c->ev_imported[b], c->stream));
    CK(c, cudaGetLastError());
    c->cur = 0;
    c->keys_valid = false;
    c->grid_valid = false;
    c->mask_valid = false;
    c->sorted_keys = nullptr;
    c->aos_stale = true;           // the records' pos / vel lanes follow at the next publish
    return NPRSPH_OK;
}

int nprsph_download_positions(nprsph_ctx* c, float* pos4, uint64_t n, uint32_t flags) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (n > c->n) return fail(c, NPRSPH_ERR_INVALID, "download larger than the particle buffer%s");
    if (n && !pos4) return fail(c, NPRSPH_ERR_INVALID, "null host buffer%s");
    if (!n) return NPRSPH_OK;
    int rc = ensure_staging(c);
    if (rc) return rc;
    if (c->d2h_pending) CK(c, cudaStreamWaitEvent(c->stream, c->ev_copied, 0));    // the previous copy has left the staging buffer
    launch_publish_positions(c->pos[c->cur], c->aos, c->stage_pos, (uint32_t)c->n, c->stream);
    CK(c, cudaEventRecord(c->ev_published, c->stream));
    CK(c, cudaStreamWaitEvent(c->d2h_stream, c->ev_published, 0));
    CK(c, cudaMemcpyAsync(pos4, c->stage_pos, n * sizeof(float4), cudaMemcpyDeviceToHost, c->d2h_stream));
    CK(c, cudaEventRecord(c->ev_copied, c->d2h_stream));
    c->d2h_pending = true;
    if (!(flags & NPRSPH_DOWNLOAD_ASYNC)) {
        CK(c, cudaStreamSynchronize(c->d2h_stream));
        c->d2h_pending = false;
    }
    return NPRSPH_OK;
}

int nprsph_walk_stats(nprsph_ctx* c, uint64_t out[5]) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (!out) return NPRSPH_ERR_INVALID;
    for (int i = 0; i < 5; i++) out[i] = 0;
    if (c->n == 0) return NPRSPH_OK;
    int rc = ensure_grid(c, true, nullptr);
    if (rc) return rc;
    unsigned long long* d_out = nullptr;
    CK(c, cudaMalloc(&d_out, 5 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d_out, 0, 5 * sizeof(unsigned long long), c->stream);
    if (e == cudaSuccess) {
        launch_walk_stats(c->pos[c->cur], c->cell_start, 0u, (uint32_t)c->n, c->grid, c->sph, d_out, c->stream);
        unsigned long long h[5];
        e = cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) for (int i = 0; i < 5; i++) out[i] = h[i];
    }
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(c, NPRSPH_ERR_CUDA, "walk_stats: %s", cudaGetErrorString(e));
    return NPRSPH_OK;
}

uint64_t nprsph_num_particles(const nprsph_ctx* c) { return c ? c->n : 0; }

// ---- pause / reset ---------------------------------------------------------------------------------
int nprsph_set_paused(nprsph_ctx* c, int paused) {
    if (!c) return NPRSPH_ERR_INVALID;
    c->paused = paused != 0;
    return NPRSPH_OK;
}
int nprsph_toggle_pause(nprsph_ctx* c) {           // 'p', Main.cpp:466-469
    if (!c) return NPRSPH_ERR_INVALID;
    c->paused = !c->paused;
    return NPRSPH_OK;
}
int nprsph_is_paused(const nprsph_ctx* c) { return c ? (c->paused ? 1 : 0) : NPRSPH_ERR_INVALID; }

int nprsph_reset(nprsph_ctx* c) {                  // 'r', Main.cpp:460-464
    GUARD(c);
    SINGLE_ONLY(c);
    return build_scene(c);                         // pause flag and constants untouched
}

// ---- stepping ----------------------------------------------------------------------------------------
static int step_once(nprsph_ctx* c) {
    int rc = ensure_grid(c, false, nullptr);
    if (rc) return rc;
    run_rho(c, false);                             // Main.cpp:295-297
    rc = run_force_integrate(c, nullptr, nullptr); // Main.cpp:298-303
    if (rc) return rc;
    c->steps_done++;
    return NPRSPH_OK;
}

// ---- a step as ONE graph launch ------------------------------------------------------------------------
// The reference's own scene is 10,000 particles: its step is ~15 kernels of 3-7 us each, i.e. bound by
// launch overhead, not by the GPU.  Once two consecutive steps were launched with identical parameters
// (same buffers, grid, constants, colliders, validity flags -- StepSig), the next one is recorded with
// stream capture and later steps replay the graph until anything in the signature changes.  The
// kernels and their order are exactly those of step_once, so results are bit-identical.
static void fill_sig(const nprsph_ctx* c, nprsph_ctx::StepSig& s) {
    memset(&s, 0, sizeof s);
    s.n = c->n; s.cap = c->cap;
    s.cur = c->cur; s.key_bits = c->key_bits; s.num_sms = c->num_sms;
    s.flags = c->cfg.flags;
    s.keys_valid = c->keys_valid; s.grid_valid = c->grid_valid; s.mask_valid = c->mask_valid;
    s.hitmask_words = c->hitmask_words;
    const void* p[18] = {c->aos, c->pos[0], c->pos[1], c->vel[0], c->vel[1], c->frc[0], c->frc[1],
                         c->keys[0], c->keys[1], c->vals[0], c->vals[1], c->counts_rho, c->counts_force,
                         c->sort_ws, c->hitmask, c->cell_start, c->gap_list, c->gap_count};
    memcpy(s.ptr, p, sizeof p);
    memcpy(&s.grid, &c->grid, sizeof s.grid);
    memcpy(&s.sph, &c->sph, sizeof s.sph);
    memcpy(&s.colliders, &c->colliders, sizeof s.colliders);
}

static int step_any(nprsph_ctx* c) {
    if (c->graph_off) return step_once(c);
    int rc = refresh_params(c);                    // the signature must see THIS step's parameters
    if (rc) return rc;
    nprsph_ctx::StepSig sig;
    fill_sig(c, sig);
    if (c->step_graph && memcmp(&sig, &c->graph_sig, sizeof sig) == 0) {
        CK(c, cudaGraphLaunch(c->step_graph, c->stream));
        const nprsph_ctx::StepPost& p = c->graph_post;
        c->cur = p.cur; c->keys_valid = p.keys_valid; c->grid_valid = p.grid_valid; c->mask_valid = p.mask_valid;
        c->aos_stale = p.aos_stale; c->sorted_keys = p.sorted_keys; c->last_perm = p.last_perm;
        c->steps_done++;
        c->graph_steps++;
        return NPRSPH_OK;
    }
    if (!(c->plain_sig_valid && memcmp(&sig, &c->plain_sig, sizeof sig) == 0)) {
        c->plain_sig = sig;                        // first step with these parameters: plain launches
        c->plain_sig_valid = true;
        return step_once(c);
    }
    // second step in a row with these parameters: record it, then run the recording
    if (c->step_graph) { cudaGraphExecDestroy(c->step_graph); c->step_graph = nullptr; }
    const nprsph_ctx::StepPost before = {c->cur, c->keys_valid, c->grid_valid, c->mask_valid, c->aos_stale, c->sorted_keys, c->last_perm};
    const uint64_t steps_before = c->steps_done;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        rc = step_once(c);
        const cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
        e = rc ? cudaErrorUnknown : e2;
        if (e == cudaSuccess) e = cudaGraphInstantiate(&c->step_graph, g, 0);
        if (g) cudaGraphDestroy(g);
    }
    if (e != cudaSuccess) {
        // nothing ran: put the host-side state back, give graphs up for this context, step plainly
        cudaGetLastError();
        c->step_graph = nullptr;
        c->graph_off = true;
        c->cur = before.cur; c->keys_valid = before.keys_valid; c->grid_valid = before.grid_valid;
        c->mask_valid = before.mask_valid; c->aos_stale = before.aos_stale;
        c->sorted_keys = before.sorted_keys; c->last_perm = before.last_perm;
        c->steps_done = steps_before;
        return step_once(c);
    }
    c->graph_sig = sig;
    c->graph_post = {c->cur, c->keys_valid, c->grid_valid, c->mask_valid, c->aos_stale, c->sorted_keys, c->last_perm};
    CK(c, cudaGraphLaunch(c->step_graph, c->stream));
    c->graph_steps++;
    return NPRSPH_OK;
}

int nprsph_step(nprsph_ctx* c, int n_steps) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (n_steps < 0) return fail(c, NPRSPH_ERR_INVALID, "negative step count%s");
    if (c->paused || c->n == 0) return NPRSPH_OK;  // if (simulate) ..., Main.cpp:293
    for (int s = 0; s < n_steps; s++) {
        int rc = step_any(c);
        if (rc) return rc;
    }
    CK(c, cudaGetLastError());
    return NPRSPH_OK;
}

int nprsph_sync(nprsph_ctx* c) {
    GUARD(c);
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->d2h_stream) CK(c, cudaStreamSynchronize(c->d2h_stream));     // asynchronous position downloads
    c->d2h_pending = false;
    return NPRSPH_OK;
}

int nprsph_pass_rho(nprsph_ctx* c) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (c->n == 0) return NPRSPH_OK;
    int rc = ensure_grid(c, true, nullptr);
    if (rc) return rc;
    run_rho(c, true);
    CK(c, cudaGetLastError());
    return NPRSPH_OK;
}

int nprsph_pass_force(nprsph_ctx* c) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (c->n == 0) return NPRSPH_OK;
    int rc = ensure_grid(c, true, nullptr);
    if (rc) return rc;
    run_force(c);
    CK(c, cudaGetLastError());
    return NPRSPH_OK;
}

int nprsph_pass_integrate(nprsph_ctx* c) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (c->n == 0) return NPRSPH_OK;
    int rc = refresh_params(c);
    if (rc) return rc;
    run_integrate(c);
    CK(c, cudaGetLastError());
    return NPRSPH_OK;
}

void* nprsph_stream(const nprsph_ctx* c) { return c ? (void*)c->stream : nullptr; }

// ---- measurement / introspection ---------------------------------------------------------------------
int nprsph_get_stats(nprsph_ctx* c, nprsph_stats* out) {
    GUARD(c);
    if (!out) return NPRSPH_ERR_INVALID;
    int rc = refresh_params(c);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    out->num_particles = c->n;
    out->steps_done = c->steps_done;
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(c->gap_count) + 1;
    launch_count_nan(c->pos[c->cur], (uint32_t)c->n, d_cnt, c->stream);
    unsigned long long h_cnt = 0;
    CK(c, cudaMemcpyAsync(&h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    out->nan_particles = h_cnt;
    out->num_cells = c->grid.num_cells;
    for (int a = 0; a < 3; a++) out->grid_dim[a] = (uint32_t)c->grid.dim[a];
    out->key_bits = (uint32_t)c->key_bits;
    out->sort_passes = (uint32_t)sort_num_passes(c->key_bits);
    out->cell_size = c->cell_size;
    out->smoothing_length = c->sph.h;
    out->paused = c->paused ? 1 : 0;
    out->cell_subdiv = effective_subdiv(c);
    out->graph_steps = c->graph_steps;
    return NPRSPH_OK;
}

int nprsph_profile_step(nprsph_ctx* c, int n_steps, float* stage_ms) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (n_steps < 1 || !stage_ms) return NPRSPH_ERR_INVALID;
    cudaEvent_t ev[8];
    for (int i = 0; i < 8; i++) CK(c, cudaEventCreate(&ev[i]));
    double acc[NPRSPH_NUM_STAGES] = {0};
    int rc = NPRSPH_OK;
    for (int s = 0; s < n_steps && rc == NPRSPH_OK && c->n; s++) {
        rc = ensure_grid(c, false, ev);            // ev[0..3]
        if (rc) break;
        CK(c, cudaEventRecord(ev[4], c->stream));
        run_rho(c, false);
        CK(c, cudaEventRecord(ev[5], c->stream));
        rc = run_force_integrate(c, nullptr, ev[6]);   // fused: everything is booked under FORCE
        if (rc) break;
        CK(c, cudaEventRecord(ev[7], c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        c->steps_done++;
        const int stage_of[7] = {NPRSPH_STAGE_KEYS, NPRSPH_STAGE_SORT, NPRSPH_STAGE_REORDER, -1,
                                 NPRSPH_STAGE_RHO, NPRSPH_STAGE_FORCE, NPRSPH_STAGE_INTEGRATE};
        for (int i = 0; i < 7; i++) {
            if (stage_of[i] < 0) continue;
            float ms = 0.f;
            CK(c, cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            acc[stage_of[i]] += ms;
        }
    }
    for (int i = 0; i < 8; i++) cudaEventDestroy(ev[i]);
    // cell table construction is fused into the reorder kernel; it is reported with REORDER
    for (int i = 0; i < NPRSPH_NUM_STAGES; i++) stage_ms[i] = (float)(acc[i] / n_steps);
    return rc;
}

int nprsph_debug_read(nprsph_ctx* c, int item, void* dst, uint64_t bytes) {
    GUARD(c);
    if (item == NPRSPH_DBG_RECORD_CTL) {           // control word of every slot pair (also in slab mode)
        const uint64_t half = (c->cap + 1) / 2;
        if (!dst || !c->hitmask || bytes != half * 4) return fail(c, NPRSPH_ERR_INVALID, "record control words: need ceil(capacity/2)*4 bytes and the column records%s");
        CK(c, cudaMemcpyAsync(dst, c->hitmask + (size_t)4 * rec_cols_of(c->grid.reach) * half, bytes, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        return NPRSPH_OK;
    }
    SINGLE_ONLY(c);
    if (!dst) return NPRSPH_ERR_INVALID;
    const void* src = nullptr;
    uint64_t need = 0;
    uint32_t* tmp = nullptr;
    switch (item) {
        case NPRSPH_DBG_SORTED_KEYS:
        case NPRSPH_DBG_CELL_START:
        case NPRSPH_DBG_LAST_PERM:
        case NPRSPH_DBG_SLOT_IDS: {
            int rc = ensure_grid(c, true, nullptr);
            if (rc) return rc;
            if (item == NPRSPH_DBG_CELL_START) { src = c->cell_start; need = ((uint64_t)c->grid.num_cells + 2) * 4; }
            else if (item == NPRSPH_DBG_SORTED_KEYS) { src = c->sorted_keys; need = c->n * 4; }
            else if (item == NPRSPH_DBG_LAST_PERM) { src = c->last_perm; need = c->n * 4; }
            else {
                CK(c, cudaMalloc(&tmp, (c->n ? c->n : 1) * 4));
                launch_slot_ids(c->pos[c->cur], tmp, (uint32_t)c->n, c->stream);
                src = tmp; need = c->n * 4;
            }
            break;
        }
        case NPRSPH_DBG_COUNTS_RHO: src = c->counts_rho; need = c->n * 4; break;
        case NPRSPH_DBG_COUNTS_FORCE: src = c->counts_force; need = c->n * 4; break;
        default: return fail(c, NPRSPH_ERR_INVALID, "unknown debug item%s");
    }
    int rc = NPRSPH_OK;
    if (need && !src) rc = fail(c, NPRSPH_ERR_STATE, "debug item not available (flag not set or no grid yet)%s");
    else if (bytes != need) rc = fail(c, NPRSPH_ERR_INVALID, "debug_read: wrong byte count%s");
    else if (need) {
        cudaError_t e = cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(c, NPRSPH_ERR_CUDA, "debug_read copy: %s", cudaGetErrorString(e));
    }
    if (tmp) { cudaStreamSynchronize(c->stream); cudaFree(tmp); }
    return rc;
}

int nprsph_sort_pairs_host(int device, const uint32_t* keys_in, const uint32_t* vals_in, uint64_t n,
                           int key_bits, uint32_t* keys_out, uint32_t* vals_out) {
    if ((n && (!keys_in || !keys_out || !vals_out)) || key_bits < 1 || key_bits > 32) return NPRSPH_ERR_INVALID;
    if (n == 0) return NPRSPH_OK;
    if (cudaSetDevice(device) != cudaSuccess) return NPRSPH_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NPRSPH_ERR_CUDA;
    uint32_t *k[2] = {nullptr, nullptr}, *v[2] = {nullptr, nullptr};
    void* ws = nullptr;
    int rc = NPRSPH_OK;
    cudaStream_t st = nullptr;
    bool in_b = false;
    cudaError_t e = cudaSuccess;
#define SK(call) do { e = (call); if (e != cudaSuccess) { rc = NPRSPH_ERR_CUDA; goto done; } } while (0)
    SK(cudaStreamCreate(&st));
    for (int b = 0; b < 2; b++) { SK(cudaMalloc(&k[b], n * 4)); SK(cudaMalloc(&v[b], n * 4)); }
    SK(cudaMalloc(&ws, sort_workspace_bytes(n)));
    SK(cudaMemcpyAsync(k[0], keys_in, n * 4, cudaMemcpyHostToDevice, st));
    if (vals_in) SK(cudaMemcpyAsync(v[0], vals_in, n * 4, cudaMemcpyHostToDevice, st));
    SK(sort_pairs(k[0], v[0], k[1], v[1], n, key_bits, vals_in == nullptr, ws, prop.multiProcessorCount, st, &in_b));
    SK(cudaMemcpyAsync(keys_out, in_b ? k[1] : k[0], n * 4, cudaMemcpyDeviceToHost, st));
    SK(cudaMemcpyAsync(vals_out, in_b ? v[1] : v[0], n * 4, cudaMemcpyDeviceToHost, st));
    SK(cudaStreamSynchronize(st));
#undef SK
done:
    if (e != cudaSuccess) g_create_error = cudaGetErrorString(e);
    for (int b = 0; b < 2; b++) { cudaFree(k[b]); cudaFree(v[b]); }
    cudaFree(ws);
    if (st) cudaStreamDestroy(st);
    return rc;
}

// ---- snapshots (the reference keeps its state on the GPU only) ---------------------
// File = 256-byte header, slot_ids[n] (uint32: original index held by each slot of the cell-ordered
// arrangement), Particle[n] (64 B records, original order).  Loading restores the arrangement as
// well, so a restarted run continues bit for bit like the uninterrupted one.
namespace {
struct SnapshotHeader {
    char magic[8];                 // "NPRSPH01"
    uint32_t header_bytes, record_bytes;
    uint64_t n, steps_done;
    nprsph_constants consts;
    nprsph_boundary bounds;
    float particle_radius, gas_const, gravity[3], damping, dt, pi;
    int32_t cell_subdiv, paused;
};
static_assert(sizeof(SnapshotHeader) <= 256, "snapshot header must fit 256 bytes");
}

int nprsph_snapshot_save(nprsph_ctx* c, const char* path) {
    GUARD(c);
    if (!path) return NPRSPH_ERR_INVALID;
    if (c->dist) return fail(c, NPRSPH_ERR_STATE, "snapshots of a slab-decomposed run are per rank: use nprsph_dist_download%s");
    int rc = publish(c);
    if (rc) return rc;
    const uint64_t n = c->n;
    nprsph_particle* rec = (nprsph_particle*)malloc((n ? n : 1) * sizeof(nprsph_particle));
    uint32_t* ids = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t* d_ids = nullptr;
    FILE* f = nullptr;
    rc = NPRSPH_OK;
    if (!rec || !ids) { rc = fail(c, NPRSPH_ERR_NOMEM, "out of host memory%s"); goto done; }
    if (n) {
        cudaError_t e = cudaMalloc(&d_ids, n * 4);
        if (e == cudaSuccess) { launch_slot_ids(c->pos[c->cur], d_ids, (uint32_t)n, c->stream); e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaMemcpyAsync(ids, d_ids, n * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(rec, c->aos, n * sizeof(nprsph_particle), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { rc = fail(c, NPRSPH_ERR_CUDA, "snapshot_save: %s", cudaGetErrorString(e)); goto done; }
    }
    {
        unsigned char raw[256];
        memset(raw, 0, sizeof raw);
        SnapshotHeader h;
        memset(&h, 0, sizeof h);
        memcpy(h.magic, "NPRSPH01", 8);
        h.header_bytes = 256; h.record_bytes = sizeof(nprsph_particle);
        h.n = n; h.steps_done = c->steps_done;
        h.consts = c->consts; h.bounds = c->bounds;
        h.particle_radius = c->cfg.particle_radius; h.gas_const = c->cfg.gas_const;
        for (int a = 0; a < 3; a++) h.gravity[a] = c->cfg.gravity[a];
        h.damping = c->cfg.damping; h.dt = c->cfg.dt; h.pi = c->cfg.pi;
        h.cell_subdiv = effective_subdiv(c); h.paused = c->paused ? 1 : 0;
        memcpy(raw, &h, sizeof h);
        f = fopen(path, "wb");
        if (!f || fwrite(raw, 1, 256, f) != 256 || fwrite(ids, 4, n, f) != n ||
            fwrite(rec, sizeof(nprsph_particle), n, f) != n)
            rc = fail(c, NPRSPH_ERR_INVALID, "cannot write snapshot %s", path);
    }
done:
    if (f) fclose(f);
    free(rec); free(ids);
    if (d_ids) cudaFree(d_ids);
    return rc;
}

int nprsph_snapshot_load(nprsph_ctx* c, const char* path) {
    GUARD(c);
    if (!path) return NPRSPH_ERR_INVALID;
    if (c->dist) return fail(c, NPRSPH_ERR_STATE, "not available in slab mode%s");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(c, NPRSPH_ERR_INVALID, "cannot open snapshot %s", path);
    unsigned char raw[256];
    SnapshotHeader h;
    nprsph_particle* rec = nullptr;
    uint32_t* ids = nullptr;
    uint32_t* d_ids = nullptr;
    int rc = NPRSPH_OK;
    if (fread(raw, 1, 256, f) != 256) { rc = fail(c, NPRSPH_ERR_INVALID, "truncated snapshot %s", path); goto done; }
    memcpy(&h, raw, sizeof h);
    if (memcmp(h.magic, "NPRSPH01", 8) != 0 || h.header_bytes != 256 || h.record_bytes != sizeof(nprsph_particle) ||
        h.n >= (1ull << 30)) { rc = fail(c, NPRSPH_ERR_INVALID, "not an NPRSPH01 snapshot: %s", path); goto done; }
    rec = (nprsph_particle*)malloc((h.n ? h.n : 1) * sizeof(nprsph_particle));
    ids = (uint32_t*)malloc((h.n ? h.n : 1) * sizeof(uint32_t));
    if (!rec || !ids) { rc = fail(c, NPRSPH_ERR_NOMEM, "out of host memory%s"); goto done; }
    if (fread(ids, 4, h.n, f) != h.n || fread(rec, sizeof(nprsph_particle), h.n, f) != h.n) {
        rc = fail(c, NPRSPH_ERR_INVALID, "truncated snapshot %s", path); goto done;
    }
    {   // the slot table must be a permutation of 0..n-1 (a duplicate would leave stale records behind)
        uint8_t* seen = (uint8_t*)calloc((h.n + 7) / 8 + 1, 1);
        if (!seen) { rc = fail(c, NPRSPH_ERR_NOMEM, "out of host memory%s"); goto done; }
        bool ok = true;
        for (uint64_t s = 0; s < h.n && ok; s++) {
            const uint32_t id = ids[s];
            ok = id < h.n && !(seen[id >> 3] & (1u << (id & 7)));
            if (ok) seen[id >> 3] |= (uint8_t)(1u << (id & 7));
        }
        free(seen);
        if (!ok) { rc = fail(c, NPRSPH_ERR_INVALID, "corrupt slot table in %s", path); goto done; }
    }
    c->consts = h.consts; c->bounds = h.bounds;
    c->cfg.particle_radius = h.particle_radius; c->cfg.gas_const = h.gas_const;
    for (int a = 0; a < 3; a++) c->cfg.gravity[a] = h.gravity[a];
    c->cfg.damping = h.damping; c->cfg.dt = h.dt; c->cfg.pi = h.pi;
    if (h.cell_subdiv >= 1 && h.cell_subdiv <= 4) c->cfg.cell_subdiv = h.cell_subdiv;
    c->paused = h.paused != 0;
    c->params_dirty = true;
    rc = nprsph_upload_particles(c, rec, h.n);        // original order: AoS view + SoA with ids
    if (rc) goto done;
    if (h.n) {                                         // re-establish the saved cell order
        const int nxt = 1 - c->cur;
        cudaError_t e = cudaMalloc(&d_ids, h.n * 4);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ids, ids, h.n * 4, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) {
            launch_gather(d_ids, c->pos[c->cur], c->vel[c->cur], c->frc[c->cur], c->pos[nxt], c->vel[nxt],
                          c->frc[nxt], (uint32_t)h.n, c->stream);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { rc = fail(c, NPRSPH_ERR_CUDA, "snapshot_load: %s", cudaGetErrorString(e)); goto done; }
        c->cur = nxt;
    }
    c->steps_done = h.steps_done;
done:
    fclose(f);
    free(rec); free(ids);
    if (d_ids) cudaFree(d_ids);
    return rc;
}

// ---- OpenGL presenter ---------------------------------------------------------------------------------
int nprsph_gl_register(nprsph_ctx* c, unsigned int gl_buffer) {
    GUARD(c);
    if (c->gl_res) return fail(c, NPRSPH_ERR_STATE, "a GL buffer is already registered%s");
    cudaError_t e = cudaGraphicsGLRegisterBuffer(&c->gl_res, gl_buffer, 0 /* cudaGraphicsRegisterFlagsNone */);
    if (e != cudaSuccess) {
        c->gl_res = nullptr;
        cudaGetLastError();      // not sticky: the caller simply has no usable GL context
        return fail(c, NPRSPH_ERR_UNSUPPORTED, "cudaGraphicsGLRegisterBuffer: %s", cudaGetErrorString(e));
    }
    return NPRSPH_OK;
}

int nprsph_gl_publish(nprsph_ctx* c) {
    GUARD(c);
    SINGLE_ONLY(c);
    if (!c->gl_res) return fail(c, NPRSPH_ERR_STATE, "no GL buffer registered%s");
    // The first publish after a register / upload / reset seeds every lane of the GL buffer from the
    // record array (the .w lanes and extras[2..3] the passes never write, Appendix B-9).  After that
    // the renderer's buffer is updated IN PLACE: the publish kernel scatters the lanes the passes
    // write (pos/vel/force .xyz, rho, pressure) from the cell-ordered state straight into the mapped
    // buffer, in original particle order (the brush pass is order-dependent, Main.cpp:355-360).
    if (!c->gl_seeded) {
        int rc = publish(c);
        if (rc) return rc;
    }
    CK(c, cudaGraphicsMapResources(1, &c->gl_res, c->stream));
    void* p = nullptr; size_t sz = 0;
    cudaError_t e = cudaGraphicsResourceGetMappedPointer(&p, &sz, c->gl_res);
    if (e == cudaSuccess && sz < c->n * sizeof(nprsph_particle)) e = cudaErrorInvalidValue;
    if (e == cudaSuccess && c->n) {
        if (!c->gl_seeded) {
            e = cudaMemcpyAsync(p, c->aos, c->n * sizeof(nprsph_particle), cudaMemcpyDeviceToDevice, c->stream);
            c->gl_seeded = e == cudaSuccess;
        } else {
            launch_publish(c->pos[c->cur], c->vel[c->cur], c->frc[c->cur], p, (uint32_t)c->n, c->stream);
            e = cudaGetLastError();
        }
    }
    cudaGraphicsUnmapResources(1, &c->gl_res, c->stream);
    if (e != cudaSuccess) return fail(c, NPRSPH_ERR_CUDA, "gl_publish: %s", cudaGetErrorString(e));
    return NPRSPH_OK;
}

int nprsph_gl_unregister(nprsph_ctx* c) {
    GUARD(c);
    if (!c->gl_res) return NPRSPH_OK;
    CK(c, cudaGraphicsUnregisterResource(c->gl_res));
    c->gl_res = nullptr;
    c->gl_seeded = false;
    return NPRSPH_OK;
}

}  // extern "C"
