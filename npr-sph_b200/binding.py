"""ctypes binding of libnprsph.so (include/nprsph.h) -- the harness side of the C ABI.

The product is the shared library; this module only loads it and mirrors its entry points
one to one so tests and bench.py read like calls into the reference's own driver
(Main.cpp: init_particles / sendUniforms / display / keyboard).  It never falls back to a
CPU implementation: if the library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (NPRSPH_LIB: an A/B build of the same library, `make VARIANT=...`; measurements only)
LIB_PATH = os.environ.get("NPRSPH_LIB") or os.path.join(_HERE, "lib", "libnprsph.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_UNSUPPORTED, ERR_COMM = -1, -2, -3, -4, -5, -6
FLAG_COUNT_NEIGHBOURS = 1
FLAG_NO_HITMASK = 2
FLAG_NO_FUSE = 4
FLAG_NO_GRAPH = 8
DOWNLOAD_ASYNC = 1

STAGES = ("keys", "sort", "cells", "reorder", "rho", "force", "integrate")
DBG_SORTED_KEYS, DBG_SLOT_IDS, DBG_CELL_START, DBG_COUNTS_RHO, DBG_COUNTS_FORCE, DBG_LAST_PERM = range(6)

# numpy view of struct Particle (Main.cpp:93-99): 16 floats per record
PARTICLE_DTYPE = np.dtype([("pos", np.float32, 4), ("vel", np.float32, 4),
                           ("force", np.float32, 4), ("extras", np.float32, 4)])


class Constants(C.Structure):       # ConstantsUniform, Main.cpp:110-116
    _fields_ = [("mass", C.c_float), ("smoothing_coeff", C.c_float), ("visc", C.c_float),
                ("resting_rho", C.c_float)]


class Boundary(C.Structure):        # BoundaryUniform, Main.cpp:118-122
    _fields_ = [("upper", C.c_float * 4), ("lower", C.c_float * 4)]


class Collider(C.Structure):        # nprsph_collider (static obstacles of the integrate pass)
    _fields_ = [("kind", C.c_uint32), ("a", C.c_float * 3), ("b", C.c_float * 3),
                ("reserved", C.c_float)]


class Slider(C.Structure):          # nprsph_slider: one widget of the "Constants Window", Main.cpp:240-247
    _fields_ = [("label", C.c_char_p), ("min", C.c_float), ("max", C.c_float), ("default", C.c_float)]


COLLIDER_SPHERE, COLLIDER_BOX = 0, 1
SLIDER_MASS, SLIDER_SMOOTHING, SLIDER_VISCOSITY, SLIDER_RESTING_DENSITY = range(4)
MAX_COLLIDERS = 8


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("stream", C.c_void_p),
                ("particle_radius", C.c_float), ("gas_const", C.c_float),
                ("gravity", C.c_float * 3), ("damping", C.c_float), ("dt", C.c_float),
                ("pi", C.c_float), ("cell_subdiv", C.c_int32), ("max_cells", C.c_uint32),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("num_particles", C.c_uint64), ("steps_done", C.c_uint64),
                ("nan_particles", C.c_uint64), ("num_cells", C.c_uint32),
                ("grid_dim", C.c_uint32 * 3), ("key_bits", C.c_uint32),
                ("sort_passes", C.c_uint32), ("cell_size", C.c_float),
                ("smoothing_length", C.c_float), ("paused", C.c_int32),
                ("cell_subdiv", C.c_int32), ("graph_steps", C.c_uint64)]


# every symbol include/nprsph.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "nprsph_abi_version": (C.c_int, []),
    "nprsph_config_default": (None, [C.POINTER(Config)]),
    "nprsph_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "nprsph_destroy": (C.c_int, [_P]),
    "nprsph_last_error": (C.c_char_p, [_P]),
    "nprsph_set_constants": (C.c_int, [_P, C.POINTER(Constants)]),
    "nprsph_get_constants": (C.c_int, [_P, C.POINTER(Constants)]),
    "nprsph_set_boundary": (C.c_int, [_P, C.POINTER(Boundary)]),
    "nprsph_slider_info": (C.c_int, [C.c_int, C.POINTER(Slider)]),
    "nprsph_set_slider": (C.c_int, [_P, C.c_int, C.c_float]),
    "nprsph_set_colliders": (C.c_int, [_P, C.POINTER(Collider), C.c_int]),
    "nprsph_get_colliders": (C.c_int, [_P, C.POINTER(Collider), C.c_int]),
    "nprsph_get_boundary": (C.c_int, [_P, C.POINTER(Boundary)]),
    "nprsph_set_config": (C.c_int, [_P, C.POINTER(Config)]),
    "nprsph_get_config": (C.c_int, [_P, C.POINTER(Config)]),
    "nprsph_scene_block": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_float,
                                     C.POINTER(C.c_float), C.c_float, C.c_uint32]),
    "nprsph_upload_particles": (C.c_int, [_P, _P, C.c_uint64]),
    "nprsph_download_particles": (C.c_int, [_P, _P, C.c_uint64]),
    "nprsph_device_particles": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "nprsph_upload_state": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "nprsph_download_positions": (C.c_int, [_P, _P, C.c_uint64, C.c_uint32]),
    "nprsph_walk_stats": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "nprsph_num_particles": (C.c_uint64, [_P]),
    "nprsph_set_paused": (C.c_int, [_P, C.c_int]),
    "nprsph_toggle_pause": (C.c_int, [_P]),
    "nprsph_is_paused": (C.c_int, [_P]),
    "nprsph_reset": (C.c_int, [_P]),
    "nprsph_step": (C.c_int, [_P, C.c_int]),
    "nprsph_sync": (C.c_int, [_P]),
    "nprsph_pass_rho": (C.c_int, [_P]),
    "nprsph_pass_force": (C.c_int, [_P]),
    "nprsph_pass_integrate": (C.c_int, [_P]),
    "nprsph_stream": (_P, [_P]),
    "nprsph_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "nprsph_profile_step": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "nprsph_debug_read": (C.c_int, [_P, C.c_int, _P, C.c_uint64]),
    "nprsph_sort_pairs_host": (C.c_int, [C.c_int, _P, _P, C.c_uint64, C.c_int, _P, _P]),
    "nprsph_snapshot_save": (C.c_int, [_P, C.c_char_p]),
    "nprsph_snapshot_load": (C.c_int, [_P, C.c_char_p]),
    "nprsph_gl_register": (C.c_int, [_P, C.c_uint]),
    "nprsph_gl_publish": (C.c_int, [_P]),
    "nprsph_gl_unregister": (C.c_int, [_P]),
    # multi-GPU (slab decomposition); the Python-side driver lives in dist.py
    "nprsph_slab_partition": (C.c_int, [C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_int32)]),
    "nprsph_slab_face_move": (C.c_int, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.c_uint32, C.c_int]),
    "nprsph_dist_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "nprsph_dist_init": (C.c_int, [_P, _P]),
    "nprsph_dist_link_local": (C.c_int, [C.POINTER(_P), C.c_int]),
    "nprsph_dist_scene_block": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_float,
                                          C.POINTER(C.c_float), C.c_float, C.c_uint32]),
    "nprsph_dist_step": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    "nprsph_dist_download": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, _P, C.c_uint64,
                                       C.POINTER(C.c_uint64)]),
    "nprsph_dist_upload": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "nprsph_dist_upload_state": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "nprsph_dist_freeze_faces": (C.c_int, [_P, C.c_int]),
    "nprsph_dist_download_positions": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, C.c_uint64,
                                                 C.POINTER(C.c_uint64), C.c_uint32]),
    "nprsph_dist_profile_step": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "nprsph_dist_get_info": (C.c_int, [_P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load libnprsph.so and bind every declared symbol.  No fallback of any kind."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C npr-sph_b200` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class NprSphError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnprsph error {code}: {msg}")
        self.code = code


def default_config(**kw) -> Config:
    cfg = Config()
    load().nprsph_config_default(C.byref(cfg))
    for k, v in kw.items():
        if k == "gravity":
            for a in range(3):
                cfg.gravity[a] = v[a]
        else:
            setattr(cfg, k, v)
    return cfg


class Simulation:
    """One libnprsph context == the reference's particle SSBO + two UBOs + three programs."""

    def __init__(self, config: Config | None = None, **cfg_kw):
        self.lib = load()
        cfg = config if config is not None else default_config(**cfg_kw)
        self._h = _P()
        rc = self.lib.nprsph_create(C.byref(cfg), C.byref(self._h))
        if rc != OK:
            raise NprSphError(rc, (self.lib.nprsph_last_error(None) or b"").decode())

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != OK:
            raise NprSphError(rc, (self.lib.nprsph_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self.lib.nprsph_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- parameters (sendUniforms, Main.cpp:274-278) ---------------------------------------
    def set_constants(self, mass=None, smoothing_coeff=None, visc=None, resting_rho=None):
        c = self.get_constants()
        for k, v in (("mass", mass), ("smoothing_coeff", smoothing_coeff), ("visc", visc),
                     ("resting_rho", resting_rho)):
            if v is not None:
                setattr(c, k, v)
        self._ck(self.lib.nprsph_set_constants(self._h, C.byref(c)))

    def get_constants(self) -> Constants:
        c = Constants()
        self._ck(self.lib.nprsph_get_constants(self._h, C.byref(c)))
        return c

    def set_boundary(self, upper, lower):
        b = Boundary()
        for a in range(3):
            b.upper[a], b.lower[a] = upper[a], lower[a]
        b.upper[3] = upper[3] if len(upper) > 3 else 1.0
        b.lower[3] = lower[3] if len(lower) > 3 else 1.0
        self._ck(self.lib.nprsph_set_boundary(self._h, C.byref(b)))

    def set_slider(self, slider_id: int, value: float):
        """One edit of a "Constants Window" slider (Main.cpp:242-245): clamped to the widget's range."""
        self._ck(self.lib.nprsph_set_slider(self._h, slider_id, value))

    def slider_info(self, slider_id: int) -> Slider:
        s = Slider()
        self._ck(self.lib.nprsph_slider_info(slider_id, C.byref(s)))
        return s

    def set_colliders(self, colliders):
        """colliders: sequence of ("sphere", centre, radius) / ("box", lower, upper) or Collider."""
        arr = (Collider * max(1, len(colliders)))()
        for i, c in enumerate(colliders):
            if isinstance(c, Collider):
                arr[i] = c
            elif c[0] == "sphere":
                arr[i] = Collider(COLLIDER_SPHERE, (C.c_float * 3)(*c[1]), (C.c_float * 3)(c[2], 0, 0), 0.0)
            else:
                arr[i] = Collider(COLLIDER_BOX, (C.c_float * 3)(*c[1]), (C.c_float * 3)(*c[2]), 0.0)
        self._ck(self.lib.nprsph_set_colliders(self._h, arr, len(colliders)))

    def get_colliders(self):
        arr = (Collider * MAX_COLLIDERS)()
        n = self.lib.nprsph_get_colliders(self._h, arr, MAX_COLLIDERS)
        if n < 0:
            self._ck(n)
        return [arr[i] for i in range(n)]

    def get_boundary(self) -> Boundary:
        b = Boundary()
        self._ck(self.lib.nprsph_get_boundary(self._h, C.byref(b)))
        return b

    def get_config(self) -> Config:
        cfg = Config()
        self._ck(self.lib.nprsph_get_config(self._h, C.byref(cfg)))
        return cfg

    def set_config(self, **kw):
        cfg = self.get_config()
        for k, v in kw.items():
            if k == "gravity":
                for a in range(3):
                    cfg.gravity[a] = v[a]
            else:
                setattr(cfg, k, v)
        self._ck(self.lib.nprsph_set_config(self._h, C.byref(cfg)))

    def apply_params(self, p):
        """Copy a parameter record with the reference's field names (mass, smoothing_coeff, ...,
        upper, lower, particle_radius, gas_const, gravity, damping, dt, pi) into this context."""
        self.set_constants(p.mass, p.smoothing_coeff, p.visc, p.resting_rho)
        self.set_boundary(list(p.upper), list(p.lower))
        self.set_config(particle_radius=p.particle_radius, gas_const=p.gas_const,
                        gravity=list(p.gravity), damping=p.damping, dt=p.dt, pi=p.pi)

    # -- particle buffer ---------------------------------------------------------------------
    def scene_block(self, nx, ny, nz, spacing=0.005, origin=None, jitter=0.0, seed=0):
        o = (C.c_float * 3)(*(origin if origin is not None else (0.0, 0.0, 0.0)))
        self._ck(self.lib.nprsph_scene_block(self._h, nx, ny, nz, spacing, o, jitter, seed))

    @property
    def num_particles(self) -> int:
        return int(self.lib.nprsph_num_particles(self._h))

    def upload(self, records: np.ndarray):
        """records: float32 array of shape (n, 16) or PARTICLE_DTYPE array of shape (n,)."""
        a = np.ascontiguousarray(records)
        assert a.nbytes % 64 == 0 and a.dtype in (np.float32, PARTICLE_DTYPE)
        self._ck(self.lib.nprsph_upload_particles(self._h, a.ctypes.data, a.nbytes // 64))

    def upload_ptr(self, host_ptr: int, n: int):
        self._ck(self.lib.nprsph_upload_particles(self._h, host_ptr, n))

    def download(self, out: np.ndarray | None = None) -> np.ndarray:
        n = self.num_particles
        if out is None:
            out = np.empty((n, 16), np.float32)
        assert out.flags.c_contiguous and out.nbytes == n * 64
        self._ck(self.lib.nprsph_download_particles(self._h, out.ctypes.data, n))
        return out

    def download_ptr(self, host_ptr: int, n: int):
        self._ck(self.lib.nprsph_download_particles(self._h, host_ptr, n))

    def upload_state(self, pos4: np.ndarray, vel4: np.ndarray):
        """The inputs of a step (positions, velocities; (n, 4) float32 each, original order)."""
        p, v = np.ascontiguousarray(pos4, np.float32), np.ascontiguousarray(vel4, np.float32)
        assert p.shape == v.shape == (self.num_particles, 4)
        self._ck(self.lib.nprsph_upload_state(self._h, p.ctypes.data, v.ctypes.data, len(p)))
        self.sync()                     # the numpy temporaries must outlive the copies

    def upload_state_ptr(self, pos_ptr: int, vel_ptr: int, n: int):
        self._ck(self.lib.nprsph_upload_state(self._h, pos_ptr, vel_ptr, n))

    def download_positions(self, out: np.ndarray | None = None) -> np.ndarray:
        """(n, 4) float32: the vec4 at offset 0 of every record, original order."""
        n = self.num_particles
        if out is None:
            out = np.empty((n, 4), np.float32)
        assert out.flags.c_contiguous and out.nbytes == n * 16
        self._ck(self.lib.nprsph_download_positions(self._h, out.ctypes.data, n, 0))
        return out

    def download_positions_ptr(self, host_ptr: int, n: int, asynchronous: bool = False):
        self._ck(self.lib.nprsph_download_positions(self._h, host_ptr, n, DOWNLOAD_ASYNC if asynchronous else 0))

    def walk_stats(self) -> dict:
        out = (C.c_uint64 * 5)()
        self._ck(self.lib.nprsph_walk_stats(self._h, out))
        return dict(zip(("distance_tests", "columns", "pair_walks", "single_walks", "neighbours"), map(int, out)))

    def device_particles(self):
        p, n = _P(), C.c_uint64()
        self._ck(self.lib.nprsph_device_particles(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # -- pause / reset (keyboard(), Main.cpp:454-476) --------------------------------------
    def set_paused(self, paused: bool):
        self._ck(self.lib.nprsph_set_paused(self._h, int(paused)))

    def toggle_pause(self):
        self._ck(self.lib.nprsph_toggle_pause(self._h))

    @property
    def paused(self) -> bool:
        return bool(self.lib.nprsph_is_paused(self._h))

    def reset(self):
        self._ck(self.lib.nprsph_reset(self._h))

    # -- stepping (display(), Main.cpp:291-305) -----------------------------------------------
    def step(self, n_steps=1):
        self._ck(self.lib.nprsph_step(self._h, n_steps))

    def sync(self):
        self._ck(self.lib.nprsph_sync(self._h))

    def pass_rho(self):
        self._ck(self.lib.nprsph_pass_rho(self._h))

    def pass_force(self):
        self._ck(self.lib.nprsph_pass_force(self._h))

    def pass_integrate(self):
        self._ck(self.lib.nprsph_pass_integrate(self._h))

    @property
    def stream(self) -> int:
        return self.lib.nprsph_stream(self._h) or 0

    # -- measurement / introspection -----------------------------------------------------------
    def stats(self) -> Stats:
        s = Stats()
        self._ck(self.lib.nprsph_get_stats(self._h, C.byref(s)))
        return s

    def save(self, path: str):
        self._ck(self.lib.nprsph_snapshot_save(self._h, str(path).encode()))

    def load(self, path: str):
        self._ck(self.lib.nprsph_snapshot_load(self._h, str(path).encode()))

    def profile_step(self, n_steps=1) -> dict:
        ms = (C.c_float * len(STAGES))()
        self._ck(self.lib.nprsph_profile_step(self._h, n_steps, ms))
        return {k: float(ms[i]) for i, k in enumerate(STAGES)}

    def debug_read(self, item: int) -> np.ndarray:
        n = self.num_particles
        count = self.stats().num_cells + 2 if item == DBG_CELL_START else n
        out = np.empty(count, np.uint32)
        self._ck(self.lib.nprsph_debug_read(self._h, item, out.ctypes.data, out.nbytes))
        return out


def sort_pairs(keys: np.ndarray, vals: np.ndarray | None, key_bits: int = 32, device: int = 0):
    """Run the hand-written onesweep sort on host arrays (stable, low key_bits bits)."""
    lib = load()
    keys = np.ascontiguousarray(keys, np.uint32)
    ko, vo = np.empty_like(keys), np.empty_like(keys)
    vp = None
    if vals is not None:
        vals = np.ascontiguousarray(vals, np.uint32)
        vp = vals.ctypes.data
    rc = lib.nprsph_sort_pairs_host(device, keys.ctypes.data, vp, len(keys), key_bits,
                                    ko.ctypes.data, vo.ctypes.data)
    if rc != OK:
        raise NprSphError(rc, (lib.nprsph_last_error(None) or b"").decode())
    return ko, vo
