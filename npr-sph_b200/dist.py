"""Harness side of the multi-GPU (slab decomposition) entry points of libnprsph.so.

Two ways to run the same protocol (include/nprsph.h, "multi-GPU"):
  * `SlabGroup.local(world, ...)`  -- all ranks are contexts of this process on ONE GPU and one
    stream (LOCAL transport): the halo / migration logic is testable on a single device;
  * `SlabGroup.nccl(rank, world, nccl_id, ...)` -- one process per GPU, ncclSend/ncclRecv between
    slab neighbours (the id comes from `unique_id()` on one rank and is broadcast by the caller,
    e.g. with torch.distributed).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import binding as B

TRANSPORT_NCCL, TRANSPORT_LOCAL = 0, 1


class DistConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("rank", C.c_int32), ("world", C.c_int32),
                ("transport", C.c_int32), ("nccl_id", C.c_uint8 * 128), ("max_own", C.c_uint64),
                ("max_ghost", C.c_uint64), ("max_migrate", C.c_uint64),
                ("rebalance_every", C.c_int32), ("reserved", C.c_int32)]


class DistInfo(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("x_begin", C.c_int32),
                ("x_end", C.c_int32), ("num_own", C.c_uint64), ("ghosts_left", C.c_uint64),
                ("ghosts_right", C.c_uint64), ("nan_particles", C.c_uint64),
                ("migrated_total", C.c_uint64), ("steps_done", C.c_uint64),
                ("cap_own", C.c_uint64), ("cap_ghost", C.c_uint64),
                ("sort_bits", C.c_uint32), ("sort_passes", C.c_uint32),
                ("rebalanced", C.c_uint64), ("last_migrated", C.c_uint32), ("reserved", C.c_uint32)]


def unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    rc = B.load().nprsph_dist_unique_id(buf)
    if rc != B.OK:
        raise B.NprSphError(rc, "nprsph_dist_unique_id failed (libnccl.so.2 missing?)")
    return bytes(buf)


def slab_partition(hist, world: int, min_width: int) -> np.ndarray:
    hist = np.ascontiguousarray(hist, np.uint64)
    bounds = np.zeros(world + 1, np.int32)
    rc = B.load().nprsph_slab_partition(hist.ctypes.data_as(C.POINTER(C.c_uint64)), len(hist), world,
                                        min_width, bounds.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != B.OK:
        raise ValueError(f"nprsph_slab_partition: {rc}")
    return bounds


SLAB_COUNTER_WORDS = 12
CNT_HALO_L, CNT_HALO_R, CNT_OWN, CNT_FREE, CNT_WIDTH, CNT_CAP_MIGRATE, CNT_COST_US = 2, 3, 6, 7, 8, 9, 10


def slab_face_move(a, b, reach: int, cap_ghost: int, by_time: bool = False) -> int:
    """The re-balancing rule for the face between the ranks whose counter blocks are a (left) and b."""
    a = np.ascontiguousarray(a, np.uint32)
    b = np.ascontiguousarray(b, np.uint32)
    assert len(a) == len(b) == SLAB_COUNTER_WORDS
    u32p = C.POINTER(C.c_uint32)
    return int(B.load().nprsph_slab_face_move(a.ctypes.data_as(u32p), b.ctypes.data_as(u32p), reach, cap_ghost,
                                              int(by_time)))


class SlabGroup:
    """The local ranks of one slab-decomposed simulation (1 with NCCL, all of them with LOCAL)."""

    def __init__(self, sims, transport):
        self.sims = sims
        self.transport = transport
        self.lib = B.load()
        self._arr = (C.c_void_p * len(sims))(*[s._h for s in sims])

    @classmethod
    def local(cls, world: int, stream: int = 0, rebalance_every: int = 0, **cfg_kw):
        own = None
        if not stream:                  # all virtual ranks must share one stream
            import torch
            own = torch.cuda.Stream()
            stream = own.cuda_stream
        sims = [B.Simulation(stream=stream, **cfg_kw) for _ in range(world)]
        sims[0]._shared_stream = own    # keep the stream object alive with the group
        for r, s in enumerate(sims):
            cfg = DistConfig(struct_size=C.sizeof(DistConfig), rank=r, world=world,
                             transport=TRANSPORT_LOCAL, rebalance_every=rebalance_every)
            s._ck(s.lib.nprsph_dist_init(s._h, C.byref(cfg)))
        g = cls(sims, TRANSPORT_LOCAL)
        rc = g.lib.nprsph_dist_link_local(g._arr, world)
        if rc != B.OK:
            raise B.NprSphError(rc, "nprsph_dist_link_local failed")
        return g

    @classmethod
    def nccl(cls, rank: int, world: int, nccl_id: bytes, max_own=0, max_ghost=0, max_migrate=0,
             rebalance_every=0, **cfg_kw):
        sim = B.Simulation(**cfg_kw)
        cfg = DistConfig(struct_size=C.sizeof(DistConfig), rank=rank, world=world,
                         transport=TRANSPORT_NCCL, max_own=max_own, max_ghost=max_ghost,
                         max_migrate=max_migrate, rebalance_every=rebalance_every)
        C.memmove(cfg.nccl_id, nccl_id, 128)
        sim._ck(sim.lib.nprsph_dist_init(sim._h, C.byref(cfg)))
        return cls([sim], TRANSPORT_NCCL)

    def _ck(self, rc):
        if rc != B.OK:
            msgs = [(s.lib.nprsph_last_error(s._h) or b"").decode() for s in self.sims]
            raise B.NprSphError(rc, " | ".join(m for m in msgs if m))

    def apply_params(self, p):
        for s in self.sims:
            s.apply_params(p)

    def set_colliders(self, colliders):
        for s in self.sims:
            s.set_colliders(colliders)

    def set_paused(self, paused: bool):
        for s in self.sims:
            s.set_paused(paused)

    def scene_block(self, nx, ny, nz, spacing=0.005, origin=None, jitter=0.0, seed=0):
        o = (C.c_float * 3)(*(origin if origin is not None else (0.0, 0.0, 0.0)))
        for s in self.sims:
            s._ck(s.lib.nprsph_dist_scene_block(s._h, nx, ny, nz, spacing, o, jitter, seed))

    def step(self, n_steps=1):
        self._ck(self.lib.nprsph_dist_step(self._arr, len(self.sims), n_steps))

    def sync(self):
        for s in self.sims:
            s.sync()

    def info(self, which=0) -> DistInfo:
        out = DistInfo()
        s = self.sims[which]
        s._ck(s.lib.nprsph_dist_get_info(s._h, C.byref(out)))
        return out

    def download(self, which=0):
        """(records (n, 16) float32, ids (n,) uint32) of the own particles of local rank `which`."""
        n = C.c_uint64()
        self._ck(self.lib.nprsph_dist_download(self._arr, len(self.sims), which, None, None, 0, C.byref(n)))
        rec = np.empty((n.value, 16), np.float32)
        ids = np.empty(n.value, np.uint32)
        self._ck(self.lib.nprsph_dist_download(self._arr, len(self.sims), which, rec.ctypes.data,
                                               ids.ctypes.data, n.value, C.byref(n)))
        return rec, ids

    def download_ptr(self, which, rec_ptr: int, ids_ptr: int, capacity: int) -> int:
        n = C.c_uint64()
        self._ck(self.lib.nprsph_dist_download(self._arr, len(self.sims), which, rec_ptr, ids_ptr,
                                               capacity, C.byref(n)))
        return n.value

    def upload_ptr(self, which, rec_ptr: int, ids_ptr: int, n: int):
        s = self.sims[which]
        s._ck(s.lib.nprsph_dist_upload(s._h, rec_ptr, ids_ptr, n))

    def freeze_faces(self, frozen: bool = True):
        """No further re-balancing decisions (for hosts that re-upload the same lists every step)."""
        for s in self.sims:
            s._ck(s.lib.nprsph_dist_freeze_faces(s._h, 1 if frozen else 0))

    def upload_state_ptr(self, which, pos_ptr: int, vel_ptr: int, n: int):
        """pos4 = (x, y, z, id bits), vel4 = (vx, vy, vz, -): the inputs of a step for local rank `which`."""
        s = self.sims[which]
        s._ck(s.lib.nprsph_dist_upload_state(s._h, pos_ptr, vel_ptr, n))

    def upload_state(self, which, pos4: np.ndarray, vel4: np.ndarray):
        p, v = np.ascontiguousarray(pos4, np.float32), np.ascontiguousarray(vel4, np.float32)
        assert p.shape == v.shape and p.shape[1] == 4
        self.upload_state_ptr(which, p.ctypes.data, v.ctypes.data, len(p))
        self.sims[which].sync()         # the numpy temporaries must outlive the copies

    def download_positions_ptr(self, which, pos_ptr: int, capacity: int, asynchronous: bool = False) -> int:
        """asynchronous: returns at once; sync() (or the next synchronous download) waits for the copy."""
        n = C.c_uint64()
        self._ck(self.lib.nprsph_dist_download_positions(self._arr, len(self.sims), which, pos_ptr,
                                                         capacity, C.byref(n), 1 if asynchronous else 0))
        return n.value

    def download_positions(self, which=0) -> np.ndarray:
        """(n, 4) float32 of local rank `which`, slot order; column 3 holds the id bits (view as uint32)."""
        n = self.download_positions_ptr(which, None, 0)
        out = np.empty((n, 4), np.float32)
        self.download_positions_ptr(which, out.ctypes.data, n)
        return out

    def profile_step(self, n_steps=1) -> dict:
        ms = (C.c_float * len(B.STAGES))()
        self._ck(self.lib.nprsph_dist_profile_step(self._arr, len(self.sims), n_steps, ms))
        return {k: float(ms[i]) for i, k in enumerate(B.STAGES)}

    def gather(self, n_global: int) -> np.ndarray:
        """All local ranks' particles placed by global index (LOCAL transport: the whole scene)."""
        out = np.full((n_global, 16), np.nan, np.float32)
        for w in range(len(self.sims)):
            rec, ids = self.download(w)
            out[ids] = rec
        return out

    def close(self):
        for s in self.sims:
            s.close()
