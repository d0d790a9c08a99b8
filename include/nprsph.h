/*
 * nprsph.h -- C ABI of libnprsph.so, the B200-native drop-in for NPR-SPH's SPH step.
 *
 * The reference has no FFI: its SPH path sits behind an implicit OpenGL-object contract
 * inside one process.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference repo's NPR-SPH/).
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Rules that mirror the reference:
 *   - one caller thread per context (same rule as the GL context thread, Main.cpp:680);
 *   - the simulation starts PAUSED (`bool simulate;` zero-initialised, Main.cpp:87) and
 *     nprsph_step() is a no-op while paused (Main.cpp:293);
 *   - parameter edits take effect at the next step (sendUniforms, Main.cpp:274-278,314);
 *   - nprsph_reset() restores the initial block and keeps the pause flag and the
 *     constants (keyboard 'r', Main.cpp:460-464);
 *   - every function returns 0 on success or a negative nprsph_status; no exceptions.
 * There is no CPU fallback: creating a context without a CUDA device fails.
 */
#ifndef NPRSPH_H
#define NPRSPH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPRSPH_ABI_VERSION 2

typedef enum nprsph_status {
    NPRSPH_OK = 0,
    NPRSPH_ERR_INVALID = -1,      /* bad argument */
    NPRSPH_ERR_CUDA = -2,         /* CUDA runtime error (sticky, see nprsph_last_error) */
    NPRSPH_ERR_NOMEM = -3,
    NPRSPH_ERR_STATE = -4,        /* call not valid in the current state */
    NPRSPH_ERR_UNSUPPORTED = -5,  /* e.g. GL interop without a current GL context */
    NPRSPH_ERR_COMM = -6          /* multi-GPU exchange failed */
} nprsph_status;

/* struct Particle -- Main.cpp:93-99; rho_pres_comp.glsl:12-18 (std430, stride 64 B).
 * extras[0] = rho, extras[1] = pressure, extras[2] ("age") and every .w lane are never
 * written by the passes and are preserved from upload to download. */
typedef struct nprsph_particle {
    float pos[4];
    float vel[4];
    float force[4];
    float extras[4];
} nprsph_particle;

/* ConstantsUniform -- Main.cpp:110-116; rho_pres_comp.glsl:25-31 (std140, binding 1, 16 B) */
typedef struct nprsph_constants {
    float mass;             /* 0.02   */
    float smoothing_coeff;  /* 4.0 ; h = smoothing_coeff * particle_radius */
    float visc;             /* 3000   */
    float resting_rho;      /* 1000   */
} nprsph_constants;

/* BoundaryUniform -- Main.cpp:118-122; integrate_comp.glsl:27-31 (binding 2, 32 B, upper first) */
typedef struct nprsph_boundary {
    float upper[4];         /* ( 0.5,  1.0,   0.5, 1) */
    float lower[4];         /* (-0.1, -0.35, -0.1, 1) */
} nprsph_boundary;

/* What the reference bakes in at compile time (Main.cpp:33-36 and the shader consts),
 * plus the knobs of the new grid pipeline.  Fill with nprsph_config_default() first. */
typedef struct nprsph_config {
    uint32_t struct_size;       /* = sizeof(nprsph_config) */
    int32_t  device;            /* CUDA device ordinal */
    void*    stream;            /* caller's cudaStream_t to run on; NULL = library-owned */
    float    particle_radius;   /* PARTICLE_RADIUS 0.005f, rho_pres_comp.glsl:5   */
    float    gas_const;         /* GAS_CONST 2000,        rho_pres_comp.glsl:33  */
    float    gravity[3];        /* G (0,-9806.65,0),      force_comp.glsl:33     */
    float    damping;           /* DAMPING 0.3,           integrate_comp.glsl:8  */
    float    dt;                /* 1.0f/NUM_PARTICLES = 1e-4, integrate_comp.glsl:33 */
    float    pi;                /* PI 3.141592741f,       rho_pres_comp.glsl:8   */
    int32_t  cell_subdiv;       /* grid cell = h/cell_subdiv (1 or 2); 0 = default */
    uint32_t max_cells;         /* cap on the cell table; 0 = default (2^28)      */
    uint32_t flags;             /* NPRSPH_FLAG_* */
    uint32_t reserved;
} nprsph_config;

#define NPRSPH_FLAG_COUNT_NEIGHBOURS 1u  /* passes also record per-particle neighbour counts */
#define NPRSPH_FLAG_NO_HITMASK 2u        /* force pass re-tests every candidate instead of using the
                                            density pass's column records (A/B measurements) */

#define NPRSPH_FLAG_NO_FUSE 4u           /* nprsph_step runs the force and integrate passes as two
                                            launches instead of the fused one (A/B measurements) */
#define NPRSPH_FLAG_NO_GRAPH 8u          /* nprsph_step launches its kernels one by one instead of replaying
                                            the CUDA graph of a step (A/B measurements; the environment
                                            variable NPRSPH_NO_GRAPH=1 does the same for every context) */

typedef struct nprsph_stats {
    uint64_t num_particles;
    uint64_t steps_done;
    uint64_t nan_particles;     /* particles whose position has a NaN component */
    uint32_t num_cells;
    uint32_t grid_dim[3];
    uint32_t key_bits;
    uint32_t sort_passes;
    float    cell_size;
    float    smoothing_length;
    int32_t  paused;
    int32_t  cell_subdiv;
    uint64_t graph_steps;       /* steps of nprsph_step that ran as one CUDA graph launch */
} nprsph_stats;

/* stage indices for nprsph_profile_step() */
enum {
    NPRSPH_STAGE_KEYS = 0,      /* stand-alone key kernel (only when keys are stale) */
    NPRSPH_STAGE_SORT,          /* histogram + onesweep digit passes */
    NPRSPH_STAGE_CELLS,         /* cell-start table */
    NPRSPH_STAGE_REORDER,       /* gather pos/vel into cell order */
    NPRSPH_STAGE_RHO,           /* rho_pres_comp.glsl */
    NPRSPH_STAGE_FORCE,         /* force_comp.glsl; in the fused step also integrate_comp.glsl */
    NPRSPH_STAGE_INTEGRATE,     /* integrate_comp.glsl (+ next-step keys); ~0 when fused into FORCE */
    NPRSPH_NUM_STAGES
};

/* items for nprsph_debug_read() (parity tests) */
enum {
    NPRSPH_DBG_SORTED_KEYS = 0, /* uint32[n]: cell key of the particle in slot s            */
    NPRSPH_DBG_SLOT_IDS,        /* uint32[n]: original particle index held in slot s        */
    NPRSPH_DBG_CELL_START,      /* uint32[num_cells+2]: first slot with key >= c            */
    NPRSPH_DBG_COUNTS_RHO,      /* uint32[n] by original index; needs FLAG_COUNT_NEIGHBOURS */
    NPRSPH_DBG_COUNTS_FORCE,    /* uint32[n] by original index                              */
    NPRSPH_DBG_LAST_PERM,       /* uint32[n]: previous slot of the particle now in slot s   */
    NPRSPH_DBG_RECORD_CTL       /* uint32[ceil(capacity/2)]: control word of every slot pair, last density pass */
};

typedef struct nprsph_ctx nprsph_ctx;

/* ---- lifetime (replaces the process-global GL objects, Main.cpp:42-56,524-538) -------- */
int  nprsph_abi_version(void);
void nprsph_config_default(nprsph_config* cfg);
int  nprsph_create(const nprsph_config* cfg, nprsph_ctx** out);
int  nprsph_destroy(nprsph_ctx* ctx);
const char* nprsph_last_error(const nprsph_ctx* ctx);   /* ctx may be NULL: create errors */

/* ---- parameters (glBufferSubData of the two UBOs, Main.cpp:274-278) ------------------- */
int nprsph_set_constants(nprsph_ctx* ctx, const nprsph_constants* c);
int nprsph_get_constants(const nprsph_ctx* ctx, nprsph_constants* c);
int nprsph_set_boundary(nprsph_ctx* ctx, const nprsph_boundary* b);
int nprsph_get_boundary(const nprsph_ctx* ctx, nprsph_boundary* b);
/* run-time version of editing the shader consts + 'r' hot reload (Main.cpp:428-450) */
int nprsph_set_config(nprsph_ctx* ctx, const nprsph_config* cfg);
int nprsph_get_config(const nprsph_ctx* ctx, nprsph_config* cfg);

/* ---- particle buffer (SSBO binding 0, Main.cpp:524-527) -------------------------------- */
/* make_grid()+init_particles(), Main.cpp:488-521: nx*ny*nz block, x outermost, z innermost,
 * pos = origin + (i,j,k)*spacing, w = 1, everything else 0.  jitter > 0 adds a seeded
 * +-jitter offset per coordinate (synthetic dam-break scenes).  Defines what reset restores. */
int nprsph_scene_block(nprsph_ctx* ctx, int nx, int ny, int nz, float spacing,
                       const float origin[3], float jitter, uint32_t seed);
/* glBufferData(..., particles.data(), ...) -- Main.cpp:526 */
int nprsph_upload_particles(nprsph_ctx* ctx, const nprsph_particle* host, uint64_t n);
/* new (the reference never reads back): records come back in ORIGINAL particle order */
int nprsph_download_particles(nprsph_ctx* ctx, nprsph_particle* host, uint64_t n);
/* device pointer to the 64-B record array in original order (what SSBO binding 0 holds);
 * valid until the next call that changes n; brought up to date by this call */
int nprsph_device_particles(nprsph_ctx* ctx, void** device_ptr, uint64_t* n);
uint64_t nprsph_num_particles(const nprsph_ctx* ctx);
/* Streaming interface (new; the reference keeps the state on the GPU for ever).
 * upload_state: the INPUTS of a step for the n existing particles, original order: pos4 / vel4 are
 *   n float4 each (xyz used; the records' .w lanes keep their values).  Force, density and pressure
 *   are outputs of the passes and read as zero until the next step.  Asynchronous on the context's
 *   stream; host buffers must stay valid until nprsph_sync() (use pinned memory to overlap).
 * download_positions: what the renderer reads -- attribute 0 = the vec4 at offset 0 of each record
 *   (Main.cpp:533-535, toon_vs.glsl:17) -- n float4 in original order, 16 B per particle instead of
 *   64.  With NPRSPH_DOWNLOAD_ASYNC the copy runs on a separate stream behind the work queued so
 *   far and the call returns at once: the result leaves while the next step's inputs arrive and
 *   the next step computes; nprsph_sync() waits for it. */
#define NPRSPH_DOWNLOAD_ASYNC 1u
int nprsph_upload_state(nprsph_ctx* ctx, const float* pos4, const float* vel4, uint64_t n);
int nprsph_download_positions(nprsph_ctx* ctx, float* pos4, uint64_t n, uint32_t flags);

/* ---- pause / reset (keyboard(), Main.cpp:454-476) ---------------------------------------- */
int nprsph_set_paused(nprsph_ctx* ctx, int paused);
int nprsph_toggle_pause(nprsph_ctx* ctx);       /* 'p' */
int nprsph_is_paused(const nprsph_ctx* ctx);
int nprsph_reset(nprsph_ctx* ctx);              /* 'r' */

/* ---- stepping (display() compute block, Main.cpp:291-305) -------------------------------- */
int nprsph_step(nprsph_ctx* ctx, int n_steps);  /* async; no-op while paused */
int nprsph_sync(nprsph_ctx* ctx);
/* one pass at a time, regardless of the pause flag (glDispatchCompute of one program):
 * rho_pres_comp.glsl / force_comp.glsl / integrate_comp.glsl */
int nprsph_pass_rho(nprsph_ctx* ctx);
/* Deviation from force_comp.glsl:59 (which reads particles[].extras[1]): the pressures p_i, p_j are
 * re-evaluated from the stored densities as max(gas_const * (rho - resting_rho), 0) with the CURRENT
 * constants -- what nprsph_pass_rho stores -- and extras[1] is rewritten with that value.  Identical
 * inside nprsph_step; differs only if the constants are edited between the two stand-alone passes or
 * the caller uploads a pressure that is not the equation of state of the uploaded density. */
int nprsph_pass_force(nprsph_ctx* ctx);
int nprsph_pass_integrate(nprsph_ctx* ctx);
void* nprsph_stream(const nprsph_ctx* ctx);     /* cudaStream_t the work is queued on */

/* ---- measurement / introspection ---------------------------------------------------------- */
int nprsph_get_stats(nprsph_ctx* ctx, nprsph_stats* out);
/* runs n_steps (ignores pause) with CUDA events around each stage on the context's stream;
 * stage_ms[NPRSPH_NUM_STAGES] receives the mean ms per step of each stage */
int nprsph_profile_step(nprsph_ctx* ctx, int n_steps, float* stage_ms);
int nprsph_debug_read(nprsph_ctx* ctx, int item, void* host_dst, uint64_t bytes);
/* work of one density pass over the current arrangement (the unit the two issue-bound neighbour
 * kernels are measured in): out = {distance tests (target x candidate), non-empty
 * columns walked, pair walks, single-target walks, neighbours found incl. self} */
int nprsph_walk_stats(nprsph_ctx* ctx, uint64_t out[5]);
/* stand-alone run of the onesweep sort on host arrays (tests): stable, low key_bits bits */
int nprsph_sort_pairs_host(int device, const uint32_t* keys_in, const uint32_t* vals_in,
                           uint64_t n, int key_bits, uint32_t* keys_out, uint32_t* vals_out);

/* ---- live parameter surface (the "Constants Window" sliders, Main.cpp:240-247) ----------------- */
enum {
    NPRSPH_SLIDER_MASS = 0,         /* ImGui::SliderFloat("Mass", ..., 0.01, 0.1)              Main.cpp:242 */
    NPRSPH_SLIDER_SMOOTHING,        /* "Smoothing" 7..10 (the default 4 lies outside)           Main.cpp:243 */
    NPRSPH_SLIDER_VISCOSITY,        /* "Viscosity" 1000..5000                                   Main.cpp:244 */
    NPRSPH_SLIDER_RESTING_DENSITY,  /* "Resting Density" 1000..5000                             Main.cpp:245 */
    NPRSPH_NUM_SLIDERS
};
typedef struct nprsph_slider { const char* label; float min, max, def; } nprsph_slider;
int nprsph_slider_info(int id, nprsph_slider* out);
/* one slider edit: clamps to the widget's range, writes the ConstantsUniform field, takes effect at
 * the next step (sendUniforms, Main.cpp:274-275) */
int nprsph_set_slider(nprsph_ctx* ctx, int id, float value);

/* ---- static colliders (new: README.md:59 "Add objects for particles to collide with") --------- */
/* Applied by the integrate pass after the Euler update and before the box walls; the response is
 * the reference's wall rule (integrate_comp.glsl:46-77): put the particle on the surface and
 * multiply the normal velocity by -damping.  Not stored in snapshots. */
#define NPRSPH_MAX_COLLIDERS 8
enum { NPRSPH_COLLIDER_SPHERE = 0,  /* a = centre, b[0] = radius                */
       NPRSPH_COLLIDER_BOX = 1 };   /* a = lower corner, b = upper corner       */
typedef struct nprsph_collider {
    uint32_t kind;
    float a[3];
    float b[3];
    float reserved;
} nprsph_collider;                  /* 32 bytes */
int nprsph_set_colliders(nprsph_ctx* ctx, const nprsph_collider* list, int n);
/* copies up to cap entries, returns the number of colliders set (negative: error) */
int nprsph_get_colliders(const nprsph_ctx* ctx, nprsph_collider* out, int cap);

/* ---- snapshots (new: the reference's state never leaves the GPU) ------------ */
/* file = 256-byte header (parameters, n, step count), the slot table of the cell-ordered
 * arrangement and n 64-byte records in original order; a loaded run continues bit for bit */
int nprsph_snapshot_save(nprsph_ctx* ctx, const char* path);
int nprsph_snapshot_load(nprsph_ctx* ctx, const char* path);

/* ---- OpenGL presenter (VAO attr 0 on the SSBO, Main.cpp:529-535) --------------------------- */
/* Registers the caller's GL buffer (>= n*64 B) through CUDA-GL interop; publish copies the
 * current records into it.  Need a current GL context; headless callers never call these. */
int nprsph_gl_register(nprsph_ctx* ctx, unsigned int gl_buffer);
int nprsph_gl_publish(nprsph_ctx* ctx);
int nprsph_gl_unregister(nprsph_ctx* ctx);

/* ---- multi-GPU: 1-D slab decomposition along x, ONE context per GPU -------------
 * No reference counterpart (single GL context, Main.cpp:668-680).  Rank r owns the particles whose
 * global x cell index is in [x_begin, x_end); every step it exchanges `reach` cell layers of ghost
 * particles with its two neighbours and hands over the particles that crossed a slab face.
 * NCCL transport: one process per GPU, ncclSend/ncclRecv on the context's stream.  LOCAL transport:
 * all ranks are contexts of one process on one stream (lets a single GPU run the whole protocol). */
#define NPRSPH_TRANSPORT_NCCL 0
#define NPRSPH_TRANSPORT_LOCAL 1

typedef struct nprsph_dist_config {
    uint32_t struct_size;       /* = sizeof(nprsph_dist_config) */
    int32_t  rank, world;
    int32_t  transport;         /* NPRSPH_TRANSPORT_* */
    uint8_t  nccl_id[128];      /* ncclUniqueId made by nprsph_dist_unique_id() on one rank */
    uint64_t max_own;           /* capacities in particles; 0 = derived from the scene */
    uint64_t max_ghost;         /* per side */
    uint64_t max_migrate;       /* per side and step */
    int32_t  rebalance_every;   /* != 0: every |rebalance_every| steps each interior slab face may move by one
                                   x cell layer towards the lighter rank; > 0 balances particle counts,
                                   < 0 the measured time of the density pass (0 = static slabs) */
    int32_t  reserved;
} nprsph_dist_config;

typedef struct nprsph_dist_info {
    int32_t  rank, world;
    int32_t  x_begin, x_end;    /* owned global x cell range */
    uint64_t num_own, ghosts_left, ghosts_right, nan_particles;
    uint64_t migrated_total, steps_done, cap_own, cap_ghost;
    uint32_t sort_bits, sort_passes;   /* key bits / digit passes of the last step's slab sort */
    uint64_t rebalanced;        /* slab-face moves of this rank so far (re-balancing) */
    uint32_t last_migrated;     /* particles handed to the neighbours in the last step */
    uint32_t reserved;
} nprsph_dist_info;

/* count-balanced slab boundaries from a per-x-plane particle histogram (pure host code):
 * bounds[0] = 0 <= ... <= bounds[world] = dimx, every slab at least min_width cells wide */
int nprsph_slab_partition(const uint64_t* hist, int dimx, int world, int min_width, int32_t* bounds);
/* Re-balancing decision for ONE slab face (pure host code, the rule nprsph_dist_step applies every
 * rebalance_every steps).  a / b = the per-step counter blocks of the rank left / right of the face,
 * NPRSPH_SLAB_COUNTER_WORDS words each: both ranks hold both blocks after the step's counter
 * exchange, so they decide alike without further communication.  Returns -1 (the face moves one x
 * cell layer to the left: a hands its last layer to b), +1 (b hands its first layer to a) or 0.
 * Faces only move with reach >= 2 (cell_subdiv >= 2): the receiver's boundary layer must hold the layer
 * handed over AND the particles that cross the old face in the same step. */
#define NPRSPH_SLAB_COUNTER_WORDS 12
enum { NPRSPH_CNT_LEAVE_L = 0, NPRSPH_CNT_LEAVE_R, NPRSPH_CNT_HALO_L, NPRSPH_CNT_HALO_R, NPRSPH_CNT_NAN,
       NPRSPH_CNT_RESERVED, NPRSPH_CNT_OWN, NPRSPH_CNT_FREE, NPRSPH_CNT_WIDTH, NPRSPH_CNT_CAP_MIGRATE,
       NPRSPH_CNT_COST_US /* smoothed duration of the rank's density pass, microseconds */ };
/* by_time != 0: weigh the ranks by NPRSPH_CNT_COST_US instead of NPRSPH_CNT_OWN (the disordered front of
 * a dam break costs more per particle than the bulk) */
int nprsph_slab_face_move(const uint32_t* a, const uint32_t* b, int reach, uint32_t cap_ghost, int by_time);
int nprsph_dist_unique_id(uint8_t id[128]);
int nprsph_dist_init(nprsph_ctx* ctx, const nprsph_dist_config* cfg);
int nprsph_dist_link_local(nprsph_ctx** ranks, int n);
/* collective: every rank passes the same GLOBAL block and keeps the particles of its slab
 * (original index = index in the global block, as in make_grid(), Main.cpp:488-505) */
int nprsph_dist_scene_block(nprsph_ctx* ctx, int nx, int ny, int nz, float spacing,
                            const float origin[3], float jitter, uint32_t seed);
/* collective step.  NCCL: ranks = {ctx}, n_local = 1.  LOCAL: all ranks in rank order. */
int nprsph_dist_step(nprsph_ctx** ranks, int n_local, int steps);
/* copies the particles ranks[which] holds (records + their global indices); every particle is held
 * by exactly one rank at any time.  Collective only right after nprsph_dist_scene_block(). */
int nprsph_dist_download(nprsph_ctx** ranks, int n_local, int which, nprsph_particle* records,
                         uint32_t* ids, uint64_t capacity, uint64_t* n_out);
/* replaces the own particles of this rank by host records + global indices (restart / e2e);
 * a record whose position lies in the slab next door is handed over by the next step */
int nprsph_dist_upload(nprsph_ctx* ctx, const nprsph_particle* records, const uint32_t* ids, uint64_t n);
/* Slab-mode counterparts of nprsph_upload_state / nprsph_download_positions (the per-step traffic
 * of a host application: 32 B in, 16 B out per particle instead of two 68-byte records), in the
 * slab's own layout: pos4[i] = (x, y, z, bit pattern of the particle's global index),
 * vel4[i] = (vx, vy, vz, ignored).  upload_state replaces the own particles of this rank like
 * nprsph_dist_upload (force / density / pressure are outputs of the next step);
 * download_positions copies the positions of the particles ranks[which] holds, in slot order
 * (pos4 NULL: only *n_out).  Same asynchrony as the single-context pair: the upload travels on a
 * copy stream into a double-buffered staging area (host buffers valid until the next
 * nprsph_dist_step returns or nprsph_sync), NPRSPH_DOWNLOAD_ASYNC lets the positions leave on a
 * third stream while the next step's inputs arrive; nprsph_sync() waits for both. */
int nprsph_dist_upload_state(nprsph_ctx* ctx, const float* pos4, const float* vel4, uint64_t n);
int nprsph_dist_download_positions(nprsph_ctx** ranks, int n_local, int which, float* pos4,
                                   uint64_t capacity, uint64_t* n_out, uint32_t flags);
/* frozen != 0: no further re-balancing decisions (a move already decided still takes effect at the
 * next step).  For hosts that upload the same particle lists again and again: an uploaded particle may
 * lie at most `reach` cell layers beyond a slab face.  Every rank of the group makes the same call. */
int nprsph_dist_freeze_faces(nprsph_ctx* ctx, int frozen);
/* collective nprsph_profile_step(): SORT = whole prepare phase, REORDER = (v, rho) halo exchange */
int nprsph_dist_profile_step(nprsph_ctx** ranks, int n_local, int steps, float* stage_ms);
int nprsph_dist_get_info(nprsph_ctx* ctx, nprsph_dist_info* out);

#ifdef __cplusplus
}
#endif
#endif /* NPRSPH_H */
