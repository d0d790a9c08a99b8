"""ctypes wrapper around oracle/_build/libsphoracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  It restates rho_pres_comp.glsl / force_comp.glsl /
integrate_comp.glsl on the CPU (see sph_oracle.c for the file:line citations).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsphoracle.so")
REC = 16


class Params(C.Structure):
    _fields_ = [
        ("mass", C.c_float), ("smoothing_coeff", C.c_float), ("visc", C.c_float),
        ("resting_rho", C.c_float),
        ("upper", C.c_float * 4), ("lower", C.c_float * 4),
        ("particle_radius", C.c_float), ("gas_const", C.c_float),
        ("gravity", C.c_float * 3), ("damping", C.c_float), ("dt", C.c_float),
        ("pi", C.c_float),
    ]


class Collider(C.Structure):
    """kind 0: sphere (a = centre, b[0] = radius); kind 1: box (a = lower, b = upper corner)."""
    _fields_ = [("kind", C.c_uint32), ("a", C.c_float * 3), ("b", C.c_float * 3),
                ("reserved", C.c_float)]


def sphere(centre, radius) -> "Collider":
    return Collider(0, (C.c_float * 3)(*centre), (C.c_float * 3)(radius, 0.0, 0.0), 0.0)


def box(lower, upper) -> "Collider":
    return Collider(1, (C.c_float * 3)(*lower), (C.c_float * 3)(*upper), 0.0)


class Grid(C.Structure):
    _fields_ = [
        ("lo", C.c_float * 3), ("inv_cell", C.c_float), ("dim", C.c_int32 * 3),
        ("reach", C.c_int32), ("num_cells", C.c_uint32), ("cell_size", C.c_float),
        ("inv_cell_d", C.c_double),
    ]


def build(force: bool = False) -> str:
    """Compile the C oracle if needed (the checker is allowed to be built anywhere)."""
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "_build/libsphoracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        u32p = C.POINTER(C.c_uint32)
        i32p = C.POINTER(C.c_int32)
        pp = C.POINTER(Params)
        gp = C.POINTER(Grid)
        L.oracle_default_params.argtypes = [pp]
        L.oracle_make_block.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, fp, fp]
        L.oracle_jitter.argtypes = [fp, C.c_int, C.c_float, C.c_uint32]
        L.oracle_smoothing_length.argtypes = [pp]
        L.oracle_smoothing_length.restype = C.c_float
        L.oracle_r2_threshold.argtypes = [C.c_float]
        L.oracle_r2_threshold.restype = C.c_float
        L.oracle_pass_rho.argtypes = [fp, C.c_int, pp, u32p]
        L.oracle_pass_force.argtypes = [fp, C.c_int, pp, u32p]
        L.oracle_pass_integrate.argtypes = [fp, C.c_int, pp]
        L.oracle_pass_integrate_colliders.argtypes = [fp, C.c_int, pp, C.POINTER(Collider), C.c_int]
        L.oracle_step.argtypes = [fp, C.c_int, pp, C.c_int]
        L.oracle_force_scale.argtypes = [fp, C.c_int, pp, fp]
        L.oracle_sample_update.argtypes = [fp, C.c_int, pp, i32p, C.c_int, fp]
        L.oracle_grid_setup.argtypes = [pp, C.c_int, C.c_uint32, gp]
        L.oracle_grid_setup.restype = C.c_int
        L.oracle_cell_keys.argtypes = [fp, C.c_int, gp, u32p]
        for name in ("oracle_pass_rho_grid", "oracle_pass_force_grid"):
            getattr(L, name).argtypes = [fp, C.c_int, pp, C.c_int, u32p]
            getattr(L, name).restype = C.c_int
        L.oracle_force_scale_grid.argtypes = [fp, C.c_int, pp, C.c_int, fp]
        L.oracle_force_scale_grid.restype = C.c_int
        L.oracle_step_grid.argtypes = [fp, C.c_int, pp, C.c_int, C.c_int]
        L.oracle_step_grid.restype = C.c_int
        L.oracle_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u32(a):
    if a is None:
        return None
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def default_params() -> Params:
    p = Params()
    lib().oracle_default_params(C.byref(p))
    return p


def dam_break_params(nx: int, ny: int, nz: int, spacing: float = 0.005) -> Params:
    """Stable dam-break recipe of SURVEY.md 8(d) for a block of nx*ny*nz particles."""
    p = default_params()
    s = np.float32(spacing)
    p.smoothing_coeff = 2.0
    p.mass = 1.2379e-4
    p.visc = 50.0
    p.gas_const = 2000.0           # the reference's own GAS_CONST (npr-sph_b200/scenes.py explains)
    p.gravity[0], p.gravity[1], p.gravity[2] = 0.0, -9.80665, 0.0
    p.dt = 1e-4
    lx, ly, lz = nx * s, ny * s, nz * s
    p.lower[0] = p.lower[1] = p.lower[2] = -s / 2
    p.upper[0], p.upper[1], p.upper[2] = 3 * lx, 2 * ly, lz + s / 2
    return p


def make_block(nx, ny, nz, spacing=0.005, origin=None) -> np.ndarray:
    P = np.empty((nx * ny * nz, REC), np.float32)
    o = None if origin is None else _fp(np.ascontiguousarray(origin, np.float32))
    lib().oracle_make_block(nx, ny, nz, spacing, o, _fp(P))
    return P


def jitter(P, amplitude, seed=1234):
    lib().oracle_jitter(_fp(P), len(P), amplitude, seed)
    return P


def smoothing_length(p) -> np.float32:
    return np.float32(lib().oracle_smoothing_length(C.byref(p)))


def r2_threshold(h) -> np.float32:
    return np.float32(lib().oracle_r2_threshold(float(h)))


def pass_rho(P, p, counts=False, grid=0):
    c = np.zeros(len(P), np.uint32) if counts else None
    if grid:
        rc = lib().oracle_pass_rho_grid(_fp(P), len(P), C.byref(p), grid, _u32(c))
        assert rc == 0, rc
    else:
        lib().oracle_pass_rho(_fp(P), len(P), C.byref(p), _u32(c))
    return c


def pass_force(P, p, counts=False, grid=0):
    c = np.zeros(len(P), np.uint32) if counts else None
    if grid:
        rc = lib().oracle_pass_force_grid(_fp(P), len(P), C.byref(p), grid, _u32(c))
        assert rc == 0, rc
    else:
        lib().oracle_pass_force(_fp(P), len(P), C.byref(p), _u32(c))
    return c


def force_scale(P, p, grid=0) -> np.ndarray:
    """Sum of |terms| of the force sums, shape (n, 3); P must hold rho/p (after pass_rho)."""
    out = np.empty((len(P), 3), np.float32)
    if grid:
        rc = lib().oracle_force_scale_grid(_fp(P), len(P), C.byref(p), grid, _fp(out))
        assert rc == 0, rc
    else:
        lib().oracle_force_scale(_fp(P), len(P), C.byref(p), _fp(out))
    return out


def pass_integrate(P, p, colliders=None):
    if colliders:
        arr = (Collider * len(colliders))(*colliders)
        lib().oracle_pass_integrate_colliders(_fp(P), len(P), C.byref(p), arr, len(colliders))
    else:
        lib().oracle_pass_integrate(_fp(P), len(P), C.byref(p))


def step(P, p, n_steps=1, grid=0):
    if grid:
        rc = lib().oracle_step_grid(_fp(P), len(P), C.byref(p), grid, n_steps)
        assert rc == 0, rc
    else:
        lib().oracle_step(_fp(P), len(P), C.byref(p), n_steps)


def sample_update(P, p, idx) -> np.ndarray:
    idx = np.ascontiguousarray(idx, np.int32)
    out = np.empty((len(idx), REC), np.float32)
    lib().oracle_sample_update(_fp(P), len(P), C.byref(p),
                               idx.ctypes.data_as(C.POINTER(C.c_int32)), len(idx), _fp(out))
    return out


def grid_setup(p, cell_subdiv=1, max_cells=0) -> Grid:
    g = Grid()
    rc = lib().oracle_grid_setup(C.byref(p), cell_subdiv, max_cells, C.byref(g))
    if rc:
        raise ValueError(f"oracle_grid_setup failed: {rc}")
    return g


def cell_keys(P, g) -> np.ndarray:
    k = np.empty(len(P), np.uint32)
    lib().oracle_cell_keys(_fp(P), len(P), C.byref(g), _u32(k))
    return k


def num_threads() -> int:
    return lib().oracle_num_threads()


# ---- GPU all-pairs oracle (tests only; needs a CUDA device) ------------------------------------
_GPU_LIB_PATH = os.path.join(_HERE, "_build", "liballpairs_gpu.so")
_gpu = None


def gpu_lib():
    global _gpu
    if _gpu is None:
        if not os.path.exists(_GPU_LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE, "_build/liballpairs_gpu.so"],
                                  stdout=subprocess.DEVNULL)
        L = C.CDLL(_GPU_LIB_PATH)
        L.oracle_gpu_pass.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(Params),
                                      C.POINTER(C.c_uint32)]
        L.oracle_gpu_pass.restype = C.c_int
        _gpu = L
    return _gpu


def gpu_pass(which: int, P, p, counts=False):
    """which: 0 = rho/pressure pass, 1 = force pass; brute force over all pairs on the GPU."""
    c = np.zeros(len(P), np.uint32) if counts else None
    rc = gpu_lib().oracle_gpu_pass(which, _fp(P), len(P), C.byref(p), _u32(c))
    assert rc == 0, "GPU all-pairs oracle failed"
    return c


def gpu_force_scale(P, p) -> np.ndarray:
    """oracle.force_scale computed by brute force on the GPU (for the 1M-particle test)."""
    out = np.zeros((len(P), 3), np.float32)
    rc = gpu_lib().oracle_gpu_pass(2, _fp(P), len(P), C.byref(p),
                                   out.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert rc == 0, "GPU all-pairs oracle failed"
    return out
