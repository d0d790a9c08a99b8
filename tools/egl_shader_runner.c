/*
 * egl_shader_runner.c -- headless runner for the reference's UNMODIFIED compute shaders
 * (BASELINE.md section 3-2): "the reference shaders on the same GPU".
 *
 *     gcc -O2 -std=c99 tools/egl_shader_runner.c -o egl_shader_runner -ldl
 *     ./egl_shader_runner /path/to/NPR-SPH [steps=100] [out.bin]
 *
 * What it does, mirroring the reference's host code step for step:
 *   - a surfaceless OpenGL 4.4+ context through EGL_EXT_platform_device (no window system);
 *   - InitShader(const char*) (InitShader.cpp:47-112): read rho_pres_comp.glsl, force_comp.glsl and
 *     integrate_comp.glsl FROM THE DIRECTORY GIVEN ON THE COMMAND LINE (nothing of the reference is
 *     stored in this repository), compile as GL_COMPUTE_SHADER, link, print the log on failure;
 *   - init_particles() / make_grid() (Main.cpp:488-539): the 10 x 100 x 10 block at spacing 0.005 as
 *     a 64-byte-per-particle SSBO at binding 0;
 *   - the two UBOs (Main.cpp:631-641) at bindings 1 and 2 filled with ConstantsData / BoundaryData
 *     (Main.cpp:110-122), as sendUniforms() does (Main.cpp:274-278);
 *   - display()'s compute block (Main.cpp:295-303): three glDispatchCompute(10, 1, 1) with
 *     glMemoryBarrier(GL_SHADER_STORAGE_BARRIER_BIT) between, `steps` times, timed with glFinish();
 *   - optionally the final particle buffer as raw 64-byte records, for comparison with the oracle
 *     (tests/golden/ holds what a correctly rounded GLSL implementation produces).
 *
 * It needs no GL or EGL headers: the few entry points and enums are declared here and resolved
 * with dlopen("libEGL.so.1") + eglGetProcAddress.  Exit status 3 = no usable EGL/GL on this machine
 * (the B200 pool's image has none: BASELINE.md section 5), 2 = bad arguments / shader build failure.
 */
#define _POSIX_C_SOURCE 200809L
#include <dlfcn.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- the slice of EGL / GL this program uses ------------------------------------------------------ */
typedef void* EGLDisplay; typedef void* EGLConfig; typedef void* EGLContext; typedef void* EGLSurface;
typedef void* EGLDeviceEXT; typedef int32_t EGLint; typedef unsigned int EGLBoolean; typedef unsigned int EGLenum;
#define EGL_PLATFORM_DEVICE_EXT 0x313F
#define EGL_OPENGL_API 0x30A2
#define EGL_SURFACE_TYPE 0x3033
#define EGL_PBUFFER_BIT 0x0001
#define EGL_RENDERABLE_TYPE 0x3040
#define EGL_OPENGL_BIT 0x0008
#define EGL_NONE 0x3038
#define EGL_CONTEXT_MAJOR_VERSION 0x3098
#define EGL_CONTEXT_MINOR_VERSION 0x30FB

typedef unsigned int GLenum, GLuint, GLbitfield; typedef int GLint, GLsizei; typedef char GLchar;
typedef ptrdiff_t GLsizeiptr, GLintptr;
#define GL_COMPUTE_SHADER 0x91B9
#define GL_COMPILE_STATUS 0x8B81
#define GL_LINK_STATUS 0x8B82
#define GL_SHADER_STORAGE_BUFFER 0x90D2
#define GL_UNIFORM_BUFFER 0x8A11
#define GL_STREAM_DRAW 0x88E0
#define GL_SHADER_STORAGE_BARRIER_BIT 0x00002000
#define GL_VERSION 0x1F02
#define GL_RENDERER 0x1F01

static void* (*p_eglGetProcAddress)(const char*);
#define GLFN(ret, name, args) static ret (*name) args
GLFN(GLuint, glCreateShader, (GLenum));
GLFN(void, glShaderSource, (GLuint, GLsizei, const GLchar* const*, const GLint*));
GLFN(void, glCompileShader, (GLuint));
GLFN(void, glGetShaderiv, (GLuint, GLenum, GLint*));
GLFN(void, glGetShaderInfoLog, (GLuint, GLsizei, GLsizei*, GLchar*));
GLFN(GLuint, glCreateProgram, (void));
GLFN(void, glAttachShader, (GLuint, GLuint));
GLFN(void, glLinkProgram, (GLuint));
GLFN(void, glGetProgramiv, (GLuint, GLenum, GLint*));
GLFN(void, glGetProgramInfoLog, (GLuint, GLsizei, GLsizei*, GLchar*));
GLFN(void, glUseProgram, (GLuint));
GLFN(void, glGenBuffers, (GLsizei, GLuint*));
GLFN(void, glBindBuffer, (GLenum, GLuint));
GLFN(void, glBufferData, (GLenum, GLsizeiptr, const void*, GLenum));
GLFN(void, glBufferSubData, (GLenum, GLintptr, GLsizeiptr, const void*));
GLFN(void, glGetBufferSubData, (GLenum, GLintptr, GLsizeiptr, void*));
GLFN(void, glBindBufferBase, (GLenum, GLuint, GLuint));
GLFN(void, glDispatchCompute, (GLuint, GLuint, GLuint));
GLFN(void, glMemoryBarrier, (GLbitfield));
GLFN(void, glFinish, (void));
GLFN(const unsigned char*, glGetString, (GLenum));

static int load_gl(void) {
#define L(name) do { *(void**)(&name) = p_eglGetProcAddress(#name); if (!name) { fprintf(stderr, "missing GL entry point %s\n", #name); return 0; } } while (0)
    L(glCreateShader); L(glShaderSource); L(glCompileShader); L(glGetShaderiv); L(glGetShaderInfoLog);
    L(glCreateProgram); L(glAttachShader); L(glLinkProgram); L(glGetProgramiv); L(glGetProgramInfoLog);
    L(glUseProgram); L(glGenBuffers); L(glBindBuffer); L(glBufferData); L(glBufferSubData); L(glGetBufferSubData);
    L(glBindBufferBase); L(glDispatchCompute); L(glMemoryBarrier); L(glFinish); L(glGetString);
#undef L
    return 1;
}

/* ---- host-side mirrors of the reference's structs (Main.cpp:93-122) and defines (:33-36) --------------- */
typedef struct { float pos[4], vel[4], force[4], extras[4]; } Particle;
typedef struct { float mass, smoothing_coeff, visc, resting_rho; } ConstantsUniform;
typedef struct { float upper[4], lower[4]; } BoundaryUniform;
enum { NUM_PARTICLES = 10000, NUM_WORK_GROUPS = 10 };
static const float PARTICLE_RADIUS = 0.005f;

/* readShaderSource + InitShader(const char*): InitShader.cpp:10-24,47-112 */
static GLuint init_shader(const char* dir, const char* file) {
    char path[4096];
    snprintf(path, sizeof path, "%s/%s", dir, file);
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", path); return 0; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    char* src = (char*)malloc((size_t)n + 1);
    if (!src || fread(src, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(src); return 0; }
    src[n] = 0; fclose(f);
    GLuint sh = glCreateShader(GL_COMPUTE_SHADER);
    const GLchar* srcs[1] = {src};
    glShaderSource(sh, 1, srcs, NULL);
    glCompileShader(sh);
    free(src);
    GLint ok = 0; char log[4096];
    glGetShaderiv(sh, GL_COMPILE_STATUS, &ok);
    if (!ok) { glGetShaderInfoLog(sh, sizeof log, NULL, log); fprintf(stderr, "%s failed to compile:\n%s\n", file, log); return 0; }
    GLuint prog = glCreateProgram();
    glAttachShader(prog, sh);
    glLinkProgram(prog);
    glGetProgramiv(prog, GL_LINK_STATUS, &ok);
    if (!ok) { glGetProgramInfoLog(prog, sizeof log, NULL, log); fprintf(stderr, "%s failed to link:\n%s\n", file, log); return 0; }
    return prog;
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <dir with the reference's *_comp.glsl> [steps] [out.bin]\n", argv[0]); return 2; }
    const char* dir = argv[1];
    const int steps = argc > 2 ? atoi(argv[2]) : 100;
    const char* out = argc > 3 ? argv[3] : NULL;

    void* egl = dlopen("libEGL.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!egl) { fprintf(stderr, "EGL unavailable: %s\n", dlerror()); return 3; }
    *(void**)(&p_eglGetProcAddress) = dlsym(egl, "eglGetProcAddress");
    EGLBoolean (*eglInitialize)(EGLDisplay, EGLint*, EGLint*); *(void**)(&eglInitialize) = dlsym(egl, "eglInitialize");
    EGLBoolean (*eglBindAPI)(EGLenum); *(void**)(&eglBindAPI) = dlsym(egl, "eglBindAPI");
    EGLBoolean (*eglChooseConfig)(EGLDisplay, const EGLint*, EGLConfig*, EGLint, EGLint*); *(void**)(&eglChooseConfig) = dlsym(egl, "eglChooseConfig");
    EGLContext (*eglCreateContext)(EGLDisplay, EGLConfig, EGLContext, const EGLint*); *(void**)(&eglCreateContext) = dlsym(egl, "eglCreateContext");
    EGLBoolean (*eglMakeCurrent)(EGLDisplay, EGLSurface, EGLSurface, EGLContext); *(void**)(&eglMakeCurrent) = dlsym(egl, "eglMakeCurrent");
    if (!p_eglGetProcAddress || !eglInitialize || !eglBindAPI || !eglChooseConfig || !eglCreateContext || !eglMakeCurrent) {
        fprintf(stderr, "EGL unavailable: libEGL.so.1 lacks the core entry points\n"); return 3;
    }
    EGLBoolean (*eglQueryDevicesEXT)(EGLint, EGLDeviceEXT*, EGLint*); *(void**)(&eglQueryDevicesEXT) = p_eglGetProcAddress("eglQueryDevicesEXT");
    EGLDisplay (*eglGetPlatformDisplayEXT)(EGLenum, void*, const EGLint*); *(void**)(&eglGetPlatformDisplayEXT) = p_eglGetProcAddress("eglGetPlatformDisplayEXT");
    if (!eglQueryDevicesEXT || !eglGetPlatformDisplayEXT) { fprintf(stderr, "EGL unavailable: no EGL_EXT_platform_device\n"); return 3; }
    EGLDeviceEXT devs[16]; EGLint ndev = 0;
    if (!eglQueryDevicesEXT(16, devs, &ndev) || ndev < 1) { fprintf(stderr, "EGL unavailable: no EGL device\n"); return 3; }
    EGLDisplay dpy = eglGetPlatformDisplayEXT(EGL_PLATFORM_DEVICE_EXT, devs[0], NULL);
    EGLint maj = 0, mnr = 0;
    if (!dpy || !eglInitialize(dpy, &maj, &mnr) || !eglBindAPI(EGL_OPENGL_API)) { fprintf(stderr, "EGL unavailable: cannot initialise a display for desktop OpenGL\n"); return 3; }
    const EGLint cfg_attr[] = {EGL_SURFACE_TYPE, EGL_PBUFFER_BIT, EGL_RENDERABLE_TYPE, EGL_OPENGL_BIT, EGL_NONE};
    EGLConfig cfg; EGLint ncfg = 0;
    if (!eglChooseConfig(dpy, cfg_attr, &cfg, 1, &ncfg) || ncfg < 1) { fprintf(stderr, "EGL unavailable: no OpenGL-capable config\n"); return 3; }
    const EGLint ctx_attr[] = {EGL_CONTEXT_MAJOR_VERSION, 4, EGL_CONTEXT_MINOR_VERSION, 4, EGL_NONE};      /* #version 440 */
    EGLContext ctx = eglCreateContext(dpy, cfg, NULL, ctx_attr);
    if (!ctx || !eglMakeCurrent(dpy, NULL, NULL, ctx)) { fprintf(stderr, "EGL unavailable: no surfaceless OpenGL 4.4 context\n"); return 3; }
    if (!load_gl()) return 3;
    printf("GL_RENDERER %s\nGL_VERSION %s\n", (const char*)glGetString(GL_RENDERER), (const char*)glGetString(GL_VERSION));

    /* reload_shader(), Main.cpp:433-450 (file names: Main.cpp:59-61) */
    GLuint prog[3];
    const char* files[3] = {"rho_pres_comp.glsl", "force_comp.glsl", "integrate_comp.glsl"};
    for (int k = 0; k < 3; k++) if (!(prog[k] = init_shader(dir, files[k]))) return 2;

    /* make_grid() + init_particles(), Main.cpp:488-527: i outermost, k innermost, w = 1 */
    Particle* P = (Particle*)calloc(NUM_PARTICLES, sizeof *P);
    if (!P) return 2;
    int q = 0;
    for (int i = 0; i < 10; i++) for (int j = 0; j < 100; j++) for (int k = 0; k < 10; k++, q++) {
        P[q].pos[0] = (float)i * PARTICLE_RADIUS; P[q].pos[1] = (float)j * PARTICLE_RADIUS;
        P[q].pos[2] = (float)k * PARTICLE_RADIUS; P[q].pos[3] = 1.0f;
    }
    GLuint ssbo, ubo_c, ubo_b;
    glGenBuffers(1, &ssbo);
    glBindBuffer(GL_SHADER_STORAGE_BUFFER, ssbo);
    glBufferData(GL_SHADER_STORAGE_BUFFER, (GLsizeiptr)(sizeof(Particle) * NUM_PARTICLES), P, GL_STREAM_DRAW);
    glBindBufferBase(GL_SHADER_STORAGE_BUFFER, 0, ssbo);
    /* UBOs, Main.cpp:631-641; contents ConstantsData / BoundaryData, Main.cpp:110-122, uploaded as in :274-278 */
    const ConstantsUniform C = {0.02f, 4.0f, 3000.0f, 1000.0f};
    const BoundaryUniform B = {{0.5f, 1.0f, 0.5f, 1.0f}, {-0.1f, -0.35f, -0.1f, 1.0f}};
    glGenBuffers(1, &ubo_c); glBindBuffer(GL_UNIFORM_BUFFER, ubo_c);
    glBufferData(GL_UNIFORM_BUFFER, sizeof C, NULL, GL_STREAM_DRAW); glBindBufferBase(GL_UNIFORM_BUFFER, 1, ubo_c);
    glBufferSubData(GL_UNIFORM_BUFFER, 0, sizeof C, &C);
    glGenBuffers(1, &ubo_b); glBindBuffer(GL_UNIFORM_BUFFER, ubo_b);
    glBufferData(GL_UNIFORM_BUFFER, sizeof B, NULL, GL_STREAM_DRAW); glBindBufferBase(GL_UNIFORM_BUFFER, 2, ubo_b);
    glBufferSubData(GL_UNIFORM_BUFFER, 0, sizeof B, &B);

    /* display(), Main.cpp:295-303: `steps` frames with the pause flag off */
    glFinish();
    const double t0 = now_s();
    for (int s = 0; s < steps; s++)
        for (int k = 0; k < 3; k++) {
            glUseProgram(prog[k]);
            glDispatchCompute(NUM_WORK_GROUPS, 1, 1);
            glMemoryBarrier(GL_SHADER_STORAGE_BARRIER_BIT);
        }
    glFinish();
    const double dt = now_s() - t0;
    printf("{\"what\": \"unmodified reference compute shaders, headless EGL\", \"particles\": %d, \"steps\": %d, "
           "\"ms_per_step\": %.4f, \"particle_updates_per_s\": %.1f}\n", NUM_PARTICLES, steps, dt / steps * 1e3,
           (double)NUM_PARTICLES * steps / dt);
    if (out) {
        glBindBuffer(GL_SHADER_STORAGE_BUFFER, ssbo);
        glGetBufferSubData(GL_SHADER_STORAGE_BUFFER, 0, (GLsizeiptr)(sizeof(Particle) * NUM_PARTICLES), P);
        FILE* f = fopen(out, "wb");
        if (!f || fwrite(P, sizeof *P, NUM_PARTICLES, f) != NUM_PARTICLES) { fprintf(stderr, "cannot write %s\n", out); return 2; }
        fclose(f);
    }
    free(P);
    return 0;
}
