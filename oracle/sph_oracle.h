/*
 * sph_oracle.h -- CPU oracle for the NPR-SPH step.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the three GLSL compute passes of the reference
 * (rho_pres_comp.glsl, force_comp.glsl, integrate_comp.glsl) plus the host-side
 * initial scene (Main.cpp:488-521).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * library (libnprsph.so) never links, loads or calls anything in oracle/.
 *
 * Parity pin: the reference ships no tests or golden vectors,
 * and its hot path is GLSL that cannot be compiled in this image.  The oracle is
 * therefore pinned against the .npz fixtures under tests/golden, which are produced by executing the
 * reference's own, unmodified shader text with the GLSL-subset SIMT interpreter in
 * tests/golden/glsl_simt.py (see tests/golden/make_golden.py).
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One record = 16 floats: pos.xyzw, vel.xyzw, force.xyzw, extras.xyzw
 * (Main.cpp:93-99; rho_pres_comp.glsl:12-18). */
#define ORACLE_REC 16

typedef struct oracle_params {
    /* ConstantsUniform, Main.cpp:110-116 */
    float mass, smoothing_coeff, visc, resting_rho;
    /* BoundaryUniform, Main.cpp:118-122 (upper first) */
    float upper[4], lower[4];
    /* shader compile-time constants made run-time */
    float particle_radius;   /* rho_pres_comp.glsl:5  */
    float gas_const;         /* rho_pres_comp.glsl:33 */
    float gravity[3];        /* force_comp.glsl:33    */
    float damping;           /* integrate_comp.glsl:8 */
    float dt;                /* integrate_comp.glsl:33 */
    float pi;                /* rho_pres_comp.glsl:8  */
} oracle_params;

/* Static collider (extension, same layout as nprsph_collider): kind 0 sphere (a = centre,
 * b[0] = radius), kind 1 box (a = lower corner, b = upper corner). */
typedef struct oracle_collider {
    uint32_t kind;
    float a[3];
    float b[3];
    float reserved;
} oracle_collider;

/* Uniform-grid definition shared (as a specification) with the CUDA path.
 * Not part of the reference; restated here so keys can be compared bit-exactly. */
typedef struct oracle_grid {
    float lo[3];
    float inv_cell;
    int32_t dim[3];
    int32_t reach;          /* cells to walk each side (= cell_subdiv) */
    uint32_t num_cells;     /* sentinel key for NaN positions == num_cells */
    float cell_size;
    double inv_cell_d;      /* 1 / cell in double: the cell coordinate is computed in fp64 */
} oracle_grid;

void oracle_default_params(oracle_params* p);
void oracle_make_block(int nx, int ny, int nz, float spacing, const float* origin,
                       float* particles);
void oracle_jitter(float* particles, int n, float amplitude, uint32_t seed);

/* derived scalars (canonical arithmetic) */
float oracle_smoothing_length(const oracle_params* p);
float oracle_r2_threshold(float h);

/* literal all-pairs passes; counts (nullable) receive per-particle neighbour counts */
void oracle_pass_rho(float* particles, int n, const oracle_params* p, uint32_t* counts);
void oracle_pass_force(float* particles, int n, const oracle_params* p, uint32_t* counts);
void oracle_pass_integrate(float* particles, int n, const oracle_params* p);
void oracle_pass_integrate_colliders(float* particles, int n, const oracle_params* p,
                                     const oracle_collider* cs, int nc);
/* sum of |terms| of the force sums per particle/component (conditioning scale for tests) */
void oracle_force_scale(const float* particles, int n, const oracle_params* p, float* scale3);
void oracle_step(float* particles, int n, const oracle_params* p, int n_steps);

/* all-pairs, but only for the m particles listed in idx (bounded CPU baseline) */
void oracle_sample_update(const float* particles, int n, const oracle_params* p,
                          const int32_t* idx, int m, float* out_records);

/* grid spec + grid-accelerated passes that are BIT-IDENTICAL to the all-pairs ones
 * (neighbours are accumulated in ascending j like the shader loop) */
int  oracle_grid_setup(const oracle_params* p, int cell_subdiv, uint32_t max_cells,
                       oracle_grid* g);
void oracle_cell_keys(const float* particles, int n, const oracle_grid* g, uint32_t* keys);
int  oracle_pass_rho_grid(float* particles, int n, const oracle_params* p, int cell_subdiv,
                          uint32_t* counts);
int  oracle_pass_force_grid(float* particles, int n, const oracle_params* p, int cell_subdiv,
                            uint32_t* counts);
int  oracle_force_scale_grid(const float* particles, int n, const oracle_params* p, int cell_subdiv,
                             float* scale3);
int  oracle_step_grid(float* particles, int n, const oracle_params* p, int cell_subdiv,
                      int n_steps);

int  oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
