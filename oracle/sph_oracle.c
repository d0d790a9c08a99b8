/*
 * sph_oracle.c -- CPU oracle for the NPR-SPH step.  TEST INFRASTRUCTURE ONLY
 * (see sph_oracle.h for who may load this).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math  (oracle/Makefile).
 * -ffp-contract=off matters: every fp32 operation below is meant to be rounded
 * individually, in the left-associative order GLSL gives the shader expressions.
 *
 * Canonical arithmetic, decided once and used by every
 * checker in this repo:
 *   h        = smoothing_coeff * PARTICLE_RADIUS                 (one fp32 multiply)
 *   pow(h,9), pow(h,6) = double pow() rounded once to fp32       (host constants)
 *   pow(q,3) = (q*q)*q ;  pow(x,2) = x*x
 *   length(d)    = sqrtf((dx*dx + dy*dy) + dz*dz)                (IEEE sqrt, no FMA)
 *   normalize(d) = d / length(d)   (IEEE divide per component; NaN when length == 0)
 *   max(a,b)     = (a < b) ? b : a                                (GLSL definition)
 * The reference's GLSL pow/normalize are driver-defined; the golden
 * vectors quantify the distance between this canon and a correctly-rounded pow.
 */
#include "sph_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Defaults: Main.cpp:110-122 (uniform blocks), :33-36 and the shader consts. */
void oracle_default_params(oracle_params* p) {
    p->mass = 0.02f;
    p->smoothing_coeff = 4.0f;
    p->visc = 3000.0f;
    p->resting_rho = 1000.0f;
    p->upper[0] = 0.5f;  p->upper[1] = 1.0f;   p->upper[2] = 0.5f;  p->upper[3] = 1.0f;
    p->lower[0] = -0.1f; p->lower[1] = -0.35f; p->lower[2] = -0.1f; p->lower[3] = 1.0f;
    p->particle_radius = 0.005f;
    p->gas_const = 2000.0f;
    p->gravity[0] = 0.0f; p->gravity[1] = -9806.65f; p->gravity[2] = 0.0f;
    p->damping = 0.3f;
    p->dt = 1.0f / 10000.0f;           /* integrate_comp.glsl:33, 1.0f / NUM_PARTICLES */
    p->pi = 3.141592741f;
}

/* make_grid() + init_particles(), Main.cpp:488-521: i outermost (x), k innermost (z);
 * pos = ((float)i * s, (float)j * s, (float)k * s, 1); vel = force = extras = 0.
 * `origin` (nullable) is added afterwards; the reference has origin 0. */
void oracle_make_block(int nx, int ny, int nz, float spacing, const float* origin,
                       float* particles) {
    size_t idx = 0;
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++)
            for (int k = 0; k < nz; k++, idx++) {
                float* r = particles + idx * ORACLE_REC;
                memset(r, 0, ORACLE_REC * sizeof(float));
                r[0] = (float)i * spacing;
                r[1] = (float)j * spacing;
                r[2] = (float)k * spacing;
                if (origin) { r[0] += origin[0]; r[1] += origin[1]; r[2] += origin[2]; }
                r[3] = 1.0f;
            }
}

static uint32_t hash32(uint32_t seed, uint32_t idx) {
    uint32_t x = seed ^ (idx * 0x9E3779B9u);
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

/* Deterministic symmetric jitter used by the synthetic dam-break scenes. */
void oracle_jitter(float* particles, int n, float amplitude, uint32_t seed) {
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            uint32_t u = hash32(seed, (uint32_t)(3 * i + a));
            float f = (float)(u >> 8) * (1.0f / 16777216.0f);
            particles[(size_t)i * ORACLE_REC + a] += (2.0f * f - 1.0f) * amplitude;
        }
}

float oracle_smoothing_length(const oracle_params* p) {
    return p->smoothing_coeff * p->particle_radius;   /* rho_pres_comp.glsl:40 */
}

static float pow_once(float x, int e) { return (float)pow((double)x, (double)e); }

/* Smallest fp32 t with sqrtf(t) >= h, so that (sqrtf(r2) < h) == (r2 < t) for all r2.
 * sqrtf is correctly rounded and monotone, which makes the two predicates identical. */
float oracle_r2_threshold(float h) {
    if (!(h > 0.0f)) return 0.0f;
    if (isinf(h)) return INFINITY;
    float t = h * h;
    while (t > 0.0f && sqrtf(nextafterf(t, 0.0f)) >= h) t = nextafterf(t, 0.0f);
    while (sqrtf(t) < h) t = nextafterf(t, INFINITY);
    return t;
}

static inline float glsl_max(float a, float b) { return (a < b) ? b : a; }

/* ---- rho_pres_comp.glsl:35-59 for one particle i over candidate list ------------ */
static inline float rho_term(const float* pi_, const float* pj, float h, float mass315,
                             float den, int* hit) {
    float dx = pi_[0] - pj[0], dy = pi_[1] - pj[1], dz = pi_[2] - pj[2];   /* :48 */
    float r = sqrtf((dx * dx + dy * dy) + dz * dz);                         /* :49 */
    if (r < h) {                                                            /* :50 */
        float q = h * h - r * r;
        float q3 = (q * q) * q;
        *hit = 1;
        return (mass315 * q3) / den;                                        /* :52 */
    }
    *hit = 0;
    return 0.0f;
}

static void rho_finish(float* rec, float rho, const oracle_params* p) {
    rec[12] = rho;                                                          /* :55 */
    rec[13] = glsl_max(p->gas_const * (rho - p->resting_rho), 0.0f);        /* :58 */
}

void oracle_pass_rho(float* P, int n, const oracle_params* p, uint32_t* counts) {
    const float h = oracle_smoothing_length(p);
    const float mass315 = p->mass * 315.0f;
    const float den = (64.0f * p->pi) * pow_once(h, 9);
    float* rho_out = (float*)malloc(sizeof(float) * (size_t)n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const float* pi_ = P + (size_t)i * ORACLE_REC;
        float rho = 0.0f;
        uint32_t c = 0;
        for (int j = 0; j < n; j++) {                                       /* :46 */
            int hit;
            float t = rho_term(pi_, P + (size_t)j * ORACLE_REC, h, mass315, den, &hit);
            if (hit) { rho += t; c++; }
        }
        rho_out[i] = rho;
        if (counts) counts[i] = c;
    }
    for (int i = 0; i < n; i++) rho_finish(P + (size_t)i * ORACLE_REC, rho_out[i], p);
    free(rho_out);
}

/* ---- force_comp.glsl:35-67 --------------------------------------------------------- */
typedef struct { float pres[3], visc[3]; } force_acc;

static inline int force_term(const float* pi_, const float* pj, float h, float mass,
                             float spiky, float lap, force_acc* a) {
    float dx = pi_[0] - pj[0], dy = pi_[1] - pj[1], dz = pi_[2] - pj[2];   /* :55 */
    float r = sqrtf((dx * dx + dy * dy) + dz * dz);                         /* :56 */
    if (!(r < h)) return 0;                                                 /* :57 */
    float rho_j = pj[12];
    float hr = h - r;
    /* :59  mass * (p_i + p_j) / (2 * rho_j) * spiky * pow(h - r, 2) * normalize(delta) */
    float s = (((mass * (pi_[13] + pj[13])) / (2.0f * rho_j)) * spiky) * (hr * hr);
    a->pres[0] -= s * (dx / r);
    a->pres[1] -= s * (dy / r);
    a->pres[2] -= s * (dz / r);
    /* :60  mass * (v_j - v_i) / rho_j * laplacian * (h - r) */
    a->visc[0] += (((mass * (pj[4] - pi_[4])) / rho_j) * lap) * hr;
    a->visc[1] += (((mass * (pj[5] - pi_[5])) / rho_j) * lap) * hr;
    a->visc[2] += (((mass * (pj[6] - pi_[6])) / rho_j) * lap) * hr;
    return 1;
}

static void force_finish(const float* rec, const force_acc* a, const oracle_params* p,
                         float* out3) {
    for (int k = 0; k < 3; k++) {
        float v = a->visc[k] * p->visc;                                     /* :63 */
        float g = rec[12] * p->gravity[k];                                  /* :65 */
        out3[k] = (a->pres[k] + v) + g;                                     /* :66 */
    }
}

void oracle_pass_force(float* P, int n, const oracle_params* p, uint32_t* counts) {
    const float h = oracle_smoothing_length(p);
    const float h6 = pow_once(h, 6);
    const float spiky = -45.0f / (p->pi * h6);                              /* :41 */
    const float lap = 45.0f / (p->pi * h6);                                 /* :42 */
    float* f_out = (float*)malloc(sizeof(float) * 3 * (size_t)n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const float* pi_ = P + (size_t)i * ORACLE_REC;
        force_acc a; memset(&a, 0, sizeof a);
        uint32_t c = 0;
        for (int j = 0; j < n; j++) {                                       /* :48 */
            if (i == j) continue;                                           /* :50-53 */
            c += (uint32_t)force_term(pi_, P + (size_t)j * ORACLE_REC, h, p->mass, spiky, lap, &a);
        }
        force_finish(pi_, &a, p, f_out + 3 * (size_t)i);
        if (counts) counts[i] = c;
    }
    for (int i = 0; i < n; i++)
        memcpy(P + (size_t)i * ORACLE_REC + 8, f_out + 3 * (size_t)i, 3 * sizeof(float));
    free(f_out);
}

/* Condition scale of the force sums: per particle and component, the sum of the absolute
 * values of everything force_comp.glsl adds up (|pressure terms| + visc*|viscosity terms| +
 * |rho*G|).  Any fp32 evaluation that re-orders or re-associates the sum can differ from
 * another by about eps * sqrt(terms) * this scale, however small the net force is, so parity
 * tests bound the element-wise error against it. */
void oracle_force_scale(const float* P, int n, const oracle_params* p, float* scale3) {
    const float h = oracle_smoothing_length(p);
    const float h6 = pow_once(h, 6);
    const float spiky = -45.0f / (p->pi * h6);
    const float lap = 45.0f / (p->pi * h6);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const float* pi_ = P + (size_t)i * ORACLE_REC;
        double acc[3] = {0, 0, 0};
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            force_acc a; memset(&a, 0, sizeof a);
            if (force_term(pi_, P + (size_t)j * ORACLE_REC, h, p->mass, spiky, lap, &a))
                for (int k = 0; k < 3; k++)
                    acc[k] += fabs((double)a.pres[k]) + fabs((double)a.visc[k] * p->visc);
        }
        for (int k = 0; k < 3; k++)
            scale3[3 * (size_t)i + k] = (float)(acc[k] + fabs((double)pi_[12] * p->gravity[k]));
    }
}

/* ---- integrate_comp.glsl:35-82 ---------------------------------------------------- */
/* Static colliders: NOT in the reference (README.md:59 lists them as future work).  This is the
 * specification the CUDA path (common.cuh:collide_sphere / collide_box) is held to bit for bit:
 * the reference's wall rule (integrate_comp.glsl:46-77) applied to the obstacle's surface. */
static void collide_one(float* r, const oracle_collider* c, float damping) {
    float* x = r; float* v = r + 4;
    if (c->kind == 0) {                                   /* sphere: a = centre, b[0] = radius */
        const float R = c->b[0];
        const float dx = x[0] - c->a[0], dy = x[1] - c->a[1], dz = x[2] - c->a[2];
        const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
        const float r2 = (xx + yy) + zz;
        const float RR = R * R;
        if (!(r2 < RR)) return;
        const float rr = sqrtf(r2);
        float n[3] = {0.0f, 1.0f, 0.0f};
        if (rr > 0.0f) { n[0] = dx / rr; n[1] = dy / rr; n[2] = dz / rr; }
        for (int k = 0; k < 3; k++) { const float t = R * n[k]; x[k] = c->a[k] + t; }
        const float t0 = v[0] * n[0], t1 = v[1] * n[1], t2 = v[2] * n[2];
        const float vn = (t0 + t1) + t2;
        const float e = 1.0f + damping;
        const float kk = e * vn;
        for (int k = 0; k < 3; k++) { const float t = kk * n[k]; v[k] = v[k] - t; }
    } else {                                              /* box: a = lower, b = upper corner */
        for (int k = 0; k < 3; k++) if (!(x[k] > c->a[k] && x[k] < c->b[k])) return;
        int best = 0; float bp = x[0] - c->a[0];
        for (int f = 1; f < 6; f++) {
            const int k = f >> 1;
            const float pen = (f & 1) ? c->b[k] - x[k] : x[k] - c->a[k];
            if (pen < bp) { bp = pen; best = f; }
        }
        const int k = best >> 1;
        x[k] = (best & 1) ? c->b[k] : c->a[k];
        v[k] = v[k] * -damping;
    }
}

static void integrate_one(float* r, const oracle_params* p, const oracle_collider* cs, int nc) {
    for (int k = 0; k < 3; k++) {
        float a = r[8 + k] / r[12];                                         /* :41 */
        float v = r[4 + k] + p->dt * a;                                     /* :42 */
        float x = r[k] + p->dt * v;                                         /* :43 */
        r[4 + k] = v;
        r[k] = x;
    }
    for (int c = 0; c < nc; c++) collide_one(r, cs + c, p->damping);
    for (int k = 0; k < 3; k++) {
        float v = r[4 + k], x = r[k];
        if (x < p->lower[k])      { x = p->lower[k]; v *= -p->damping; }    /* :46-77 */
        else if (x > p->upper[k]) { x = p->upper[k]; v *= -p->damping; }
        r[4 + k] = v;                                                       /* :80 */
        r[k] = x;                                                           /* :81 */
    }
}

void oracle_pass_integrate(float* P, int n, const oracle_params* p) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) integrate_one(P + (size_t)i * ORACLE_REC, p, NULL, 0);
}

void oracle_pass_integrate_colliders(float* P, int n, const oracle_params* p,
                                     const oracle_collider* cs, int nc) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) integrate_one(P + (size_t)i * ORACLE_REC, p, cs, nc);
}

/* display() compute block, Main.cpp:293-304: rho -> barrier -> force -> barrier -> integrate */
void oracle_step(float* P, int n, const oracle_params* p, int n_steps) {
    for (int s = 0; s < n_steps; s++) {
        oracle_pass_rho(P, n, p, NULL);
        oracle_pass_force(P, n, p, NULL);
        oracle_pass_integrate(P, n, p);
    }
}

/* Bounded-sample CPU baseline: the reference algorithm (all-pairs j loop) for the m
 * particles in idx only.  extras of all particles must already hold rho/p (the force
 * loop reads rho_j, p_j).  Writes m updated records to out_records. */
void oracle_sample_update(const float* P, int n, const oracle_params* p, const int32_t* idx,
                          int m, float* out) {
    const float h = oracle_smoothing_length(p);
    const float mass315 = p->mass * 315.0f;
    const float den = (64.0f * p->pi) * pow_once(h, 9);
    const float h6 = pow_once(h, 6);
    const float spiky = -45.0f / (p->pi * h6);
    const float lap = 45.0f / (p->pi * h6);
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < m; s++) {
        const int i = idx[s];
        float* rec = out + (size_t)s * ORACLE_REC;
        memcpy(rec, P + (size_t)i * ORACLE_REC, ORACLE_REC * sizeof(float));
        float rho = 0.0f;
        for (int j = 0; j < n; j++) {
            int hit;
            float t = rho_term(rec, P + (size_t)j * ORACLE_REC, h, mass315, den, &hit);
            if (hit) rho += t;
        }
        rho_finish(rec, rho, p);
        force_acc a; memset(&a, 0, sizeof a);
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            force_term(rec, P + (size_t)j * ORACLE_REC, h, p->mass, spiky, lap, &a);
        }
        force_finish(rec, &a, p, rec + 8);
        integrate_one(rec, p, NULL, 0);
    }
}

/* ==== uniform grid (specification shared with the CUDA path; DESIGN.md "grid") ====== */

int oracle_grid_setup(const oracle_params* p, int k, uint32_t max_cells, oracle_grid* g) {
    const float h = oracle_smoothing_length(p);
    if (!(h > 0.0f) || isinf(h) || k < 1 || k > 4) return -1;
    double ext[3];
    for (int a = 0; a < 3; a++) {
        double e = (double)p->upper[a] - (double)p->lower[a];
        if (!(e == e) || isinf(e)) return -2;
        ext[a] = e > 0.0 ? e : 0.0;
    }
    if (max_cells == 0) max_cells = 1u << 28;
    /* cell = base * (1 + widen), widen = 2^-14: two particles closer than h (fp32 predicate, rounding
     * <= 2^-22 relative) are fewer than `reach` cells apart, and the cell coordinate is computed in
     * fp64 (error ~ dim * 2^-52 cells), so `reach` cells suffice for any grid size
     * (DESIGN.md "Grid") */
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
        double total = dims[0] * dims[1] * dims[2];
        if (total > (double)max_cells) { base *= cbrt(total / (double)max_cells) * 1.0001; continue; }
        break;
    }
    for (int a = 0; a < 3; a++) { g->lo[a] = p->lower[a]; g->dim[a] = (int32_t)dims[a]; }
    g->inv_cell_d = 1.0 / cell;
    g->inv_cell = (float)g->inv_cell_d;
    g->cell_size = (float)cell;
    g->reach = k;
    g->num_cells = (uint32_t)(g->dim[0] * (int64_t)g->dim[1] * g->dim[2]);
    return 0;
}

static inline int cell_coord(float x, float lo, double inv_cell, int dim) {
    double d = (double)x - (double)lo;
    double u = d * inv_cell;
    if (!(u >= 0.0)) u = 0.0;
    double top = (double)(dim - 1);
    if (u > top) u = top;
    return (int)u;
}

static inline uint32_t cell_key(const float* pos, const oracle_grid* g) {
    if (pos[0] != pos[0] || pos[1] != pos[1] || pos[2] != pos[2]) return g->num_cells;
    int cx = cell_coord(pos[0], g->lo[0], g->inv_cell_d, g->dim[0]);
    int cy = cell_coord(pos[1], g->lo[1], g->inv_cell_d, g->dim[1]);
    int cz = cell_coord(pos[2], g->lo[2], g->inv_cell_d, g->dim[2]);
    return ((uint32_t)cx * (uint32_t)g->dim[1] + (uint32_t)cy) * (uint32_t)g->dim[2] + (uint32_t)cz;
}

void oracle_cell_keys(const float* P, int n, const oracle_grid* g, uint32_t* keys) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) keys[i] = cell_key(P + (size_t)i * ORACLE_REC, g);
}

typedef struct { uint32_t* start; int32_t* items; uint32_t* keys; } cell_lists;

static int build_lists(const float* P, int n, const oracle_grid* g, cell_lists* L) {
    uint32_t nc = g->num_cells + 1;   /* + sentinel cell */
    L->keys = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    L->start = (uint32_t*)calloc((size_t)nc + 1, sizeof(uint32_t));
    L->items = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    if (!L->keys || !L->start || !L->items) return -1;
    oracle_cell_keys(P, n, g, L->keys);
    for (int i = 0; i < n; i++) L->start[L->keys[i] + 1]++;
    for (uint32_t c = 0; c < nc; c++) L->start[c + 1] += L->start[c];
    uint32_t* fill = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nc);
    memcpy(fill, L->start, sizeof(uint32_t) * (size_t)nc);
    for (int i = 0; i < n; i++) L->items[fill[L->keys[i]]++] = i;   /* ascending i per cell */
    free(fill);
    return 0;
}

static void free_lists(cell_lists* L) { free(L->keys); free(L->start); free(L->items); }

static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

/* Collect every j whose cell is within `reach` cells of i's cell, in ascending j. */
static int gather_candidates(int i, const cell_lists* L, const oracle_grid* g, int32_t** buf,
                             int* cap) {
    uint32_t key = L->keys[i];
    if (key >= g->num_cells) return 0;
    int cz = (int)(key % (uint32_t)g->dim[2]);
    int cy = (int)((key / (uint32_t)g->dim[2]) % (uint32_t)g->dim[1]);
    int cx = (int)(key / ((uint32_t)g->dim[2] * (uint32_t)g->dim[1]));
    int m = 0;
    for (int x = cx - g->reach; x <= cx + g->reach; x++) {
        if (x < 0 || x >= g->dim[0]) continue;
        for (int y = cy - g->reach; y <= cy + g->reach; y++) {
            if (y < 0 || y >= g->dim[1]) continue;
            for (int z = cz - g->reach; z <= cz + g->reach; z++) {
                if (z < 0 || z >= g->dim[2]) continue;
                uint32_t c = ((uint32_t)x * (uint32_t)g->dim[1] + (uint32_t)y) * (uint32_t)g->dim[2] + (uint32_t)z;
                for (uint32_t s = L->start[c]; s < L->start[c + 1]; s++) {
                    if (m == *cap) { *cap = *cap ? *cap * 2 : 1024; *buf = (int32_t*)realloc(*buf, sizeof(int32_t) * (size_t)*cap); }
                    (*buf)[m++] = L->items[s];
                }
            }
        }
    }
    qsort(*buf, (size_t)m, sizeof(int32_t), cmp_i32);
    return m;
}

int oracle_pass_rho_grid(float* P, int n, const oracle_params* p, int k, uint32_t* counts) {
    oracle_grid g; cell_lists L;
    if (oracle_grid_setup(p, k, 0, &g)) return -1;
    if (build_lists(P, n, &g, &L)) return -2;
    const float h = oracle_smoothing_length(p);
    const float mass315 = p->mass * 315.0f;
    const float den = (64.0f * p->pi) * pow_once(h, 9);
    float* rho_out = (float*)malloc(sizeof(float) * (size_t)n);
#pragma omp parallel
    {
        int32_t* buf = NULL; int cap = 0;
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            const float* pi_ = P + (size_t)i * ORACLE_REC;
            int m = gather_candidates(i, &L, &g, &buf, &cap);
            float rho = 0.0f; uint32_t c = 0;
            for (int s = 0; s < m; s++) {
                int hit;
                float t = rho_term(pi_, P + (size_t)buf[s] * ORACLE_REC, h, mass315, den, &hit);
                if (hit) { rho += t; c++; }
            }
            rho_out[i] = rho;
            if (counts) counts[i] = c;
        }
        free(buf);
    }
    for (int i = 0; i < n; i++) rho_finish(P + (size_t)i * ORACLE_REC, rho_out[i], p);
    free(rho_out); free_lists(&L);
    return 0;
}

int oracle_pass_force_grid(float* P, int n, const oracle_params* p, int k, uint32_t* counts) {
    oracle_grid g; cell_lists L;
    if (oracle_grid_setup(p, k, 0, &g)) return -1;
    if (build_lists(P, n, &g, &L)) return -2;
    const float h = oracle_smoothing_length(p);
    const float h6 = pow_once(h, 6);
    const float spiky = -45.0f / (p->pi * h6);
    const float lap = 45.0f / (p->pi * h6);
    float* f_out = (float*)malloc(sizeof(float) * 3 * (size_t)n);
#pragma omp parallel
    {
        int32_t* buf = NULL; int cap = 0;
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            const float* pi_ = P + (size_t)i * ORACLE_REC;
            int m = gather_candidates(i, &L, &g, &buf, &cap);
            force_acc a; memset(&a, 0, sizeof a);
            uint32_t c = 0;
            for (int s = 0; s < m; s++) {
                if (buf[s] == i) continue;
                c += (uint32_t)force_term(pi_, P + (size_t)buf[s] * ORACLE_REC, h, p->mass, spiky, lap, &a);
            }
            force_finish(pi_, &a, p, f_out + 3 * (size_t)i);
            if (counts) counts[i] = c;
        }
        free(buf);
    }
    for (int i = 0; i < n; i++)
        memcpy(P + (size_t)i * ORACLE_REC + 8, f_out + 3 * (size_t)i, 3 * sizeof(float));
    free(f_out); free_lists(&L);
    return 0;
}

/* oracle_force_scale over the grid's candidate lists (same terms: a candidate that fails the
 * predicate contributes nothing), for scenes too large for the all-pairs loop. */
int oracle_force_scale_grid(const float* P, int n, const oracle_params* p, int k, float* scale3) {
    oracle_grid g; cell_lists L;
    if (oracle_grid_setup(p, k, 0, &g)) return -1;
    if (build_lists(P, n, &g, &L)) return -2;
    const float h = oracle_smoothing_length(p);
    const float h6 = pow_once(h, 6);
    const float spiky = -45.0f / (p->pi * h6);
    const float lap = 45.0f / (p->pi * h6);
#pragma omp parallel
    {
        int32_t* buf = NULL; int cap = 0;
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            const float* pi_ = P + (size_t)i * ORACLE_REC;
            int m = gather_candidates(i, &L, &g, &buf, &cap);
            double acc[3] = {0, 0, 0};
            for (int s = 0; s < m; s++) {
                if (buf[s] == i) continue;
                force_acc a; memset(&a, 0, sizeof a);
                if (force_term(pi_, P + (size_t)buf[s] * ORACLE_REC, h, p->mass, spiky, lap, &a))
                    for (int q = 0; q < 3; q++)
                        acc[q] += fabs((double)a.pres[q]) + fabs((double)a.visc[q] * p->visc);
            }
            for (int q = 0; q < 3; q++)
                scale3[3 * (size_t)i + q] = (float)(acc[q] + fabs((double)pi_[12] * p->gravity[q]));
        }
        free(buf);
    }
    free_lists(&L);
    return 0;
}

int oracle_step_grid(float* P, int n, const oracle_params* p, int k, int n_steps) {
    for (int s = 0; s < n_steps; s++) {
        int rc = oracle_pass_rho_grid(P, n, p, k, NULL);
        if (rc) return rc;
        rc = oracle_pass_force_grid(P, n, p, k, NULL);
        if (rc) return rc;
        oracle_pass_integrate(P, n, p);
    }
    return 0;
}
