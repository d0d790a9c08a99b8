#!/usr/bin/env python
"""bench.py -- SPH particle-updates/s of libnprsph on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # reference algorithm on host cores

One "step" = one pass of the hot path (grid maintenance + rho/pressure + force + integrate,
i.e. display()'s compute block, Main.cpp:291-305) over every particle of the workload.

Workload (N = 1): the 16,777,216-particle fp32 dam break (BASELINE.json configs[2], the
configuration the HBM-roofline headline is quoted on), 256^3 lattice block, seeded jitter, timed
AFTER the dam has broken (--evolve-steps, default 2000): a real, disordered fluid.  The same
kernels on the still-standing lattice (the friendliest arrangement) are reported under "lattice".
N > 1: ONE global dam break of 16 Mi particles per GPU split into x slabs (weak scaling), after a
parity check of the NCCL path against a single-GPU run ("dist_parity"); BASELINE configs[3] (64 Mi
strong scaling) and configs[4] (32 Mi per GPU, migration every step) are timed in the same run
under "configs".  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SPH particle-updates/s (rho+force+integrate)"
UNIT = "particle-updates/s"
SPACING = 0.005
JITTER, SEED = 1e-4 * SPACING, 1234
# algorithmic (compulsory) bytes per particle per launch, SURVEY.md 8(d) / DESIGN.md 4.  The fused
# force+integrate launch reads pos+vel+(rho,p) once (48) and writes force (16) and pos+vel (32).
ALGO_BYTES = {"rho": 32, "force": 64, "integrate": 96}
ALGO_BYTES_FUSED = {"rho": 32, "force": 96}
KERNEL_OF_STAGE = {"rho": "k_rho", "force": "k_force_records", "integrate": "k_integrate",
                   "sort": "k_onesweep(+k_radix_hist)", "reorder": "k_reorder_cells(+k_fill_gaps)",
                   "keys": "k_keys"}
# SASS-counted issue slots of the density pass per distance test (one candidate against one target):
# 40 instructions per two candidates per target pair (profiles/r2_sass_k_rho_loop.txt)
RHO_INSTR_PER_TEST = 10.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=256, help="lattice block is side^3 particles per GPU")
    ap.add_argument("--subdiv", type=int, default=0, help="grid cells per smoothing length (0 = automatic: 2 here)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--e2e-overlap", action="store_true",
                    help="N > 1: asynchronous position downloads in the e2e leg (N = 1 always overlaps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--evolve-steps", type=int, default=2000,
                    help="steps the dam break runs before the timed region (0 = time the standing lattice)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1 main line: weak = side^3 per GPU; strong = BASELINE configs[3], 512x256x512 in total")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the secondary BASELINE configs[3]/[4] sections")
    ap.add_argument("--extra-evolve-steps", type=int, default=1200)
    ap.add_argument("--rebalance-every", type=int, default=8,
                    help="N > 1: slab faces may move one x layer every so many steps (0 = static slabs)")
    ap.add_argument("--balance", default="time", choices=["time", "count"],
                    help="N > 1: what re-balancing equalises: the measured time of the density pass or particle counts")
    return ap.parse_args()


def workload_name(dims, evolve_steps, n_gpus):
    nx, ny, nz = dims
    state = f"timed after {evolve_steps} steps of dam break" if evolve_steps else "standing lattice"
    return (f"dam_break_fp32 block {nx}x{ny}x{nz} = {nx * ny * nz} particles on {n_gpus} GPU(s), h=2s, "
            f"recipe npr-sph_b200/scenes.py (gas_const 2000), jitter seed {SEED}, {state}")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML (the source nvidia-smi
    itself reads) is polled every 20 ms from a thread; `nvidia-smi -lms` is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.samples, self.marks = index, None, [], []
        self.stop_flag = threading.Event()
        self.source = None

    def _nvml_loop(self, nv, handle, max_mhz):
        R = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
             "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((time.time(), float(mhz), float(max_mhz),
                                     [k for k, b in R.items() if bits & b]))
            except Exception:
                pass
            time.sleep(0.02)

    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append((time.time(), float(f[0]), float(f[1]),
                                     [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.source = "nvml"
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h, mx), daemon=True)
            self.t.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._smi_loop, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def mark(self):
        """Wall-clock marker: call at the start and at the end of the timed region."""
        self.marks.append(time.time())

    def stop(self) -> dict:
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        time.sleep(0.03)
        samples, window = self.samples, "sampler lifetime (warm-up + timed region, same workload)"
        if len(self.marks) == 2:
            inside = [x for x in samples if self.marks[0] <= x[0] <= self.marks[1]]
            if len(inside) >= 2:
                samples, window = inside, "timed region"
        sm = [x[1] for x in samples]
        reasons = sorted({r for x in samples for r in x[3]})
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max((x[2] for x in samples), default=None), "samples": len(sm),
                "window": window, "source": self.source, "reasons": reasons}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def relaunch_under_torchrun(args):
    """`python bench.py --gpus N` (N>1) typed by hand: start one rank per GPU ourselves."""
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.abspath(__file__)] + sys.argv[1:]
    sys.exit(subprocess.call(cmd))


# ------------------------------------------------------------------------------------------------
class SingleRunner:
    """N = 1: one context, the plain C-ABI calls (nprsph_step / upload_state / download_positions)."""

    def __init__(self, sph, dims, args, local, stream, flags=0):
        self.dims = dims
        self.n_own = self.n_total = dims[0] * dims[1] * dims[2]
        self.p = sph.scenes.dam_break_params(*dims)
        # (max_cells: the 64 Mi box has 4.0e8 cells of h/2; the default cap of 2^28 would enlarge them)
        self.sim = sph.Simulation(device=local, stream=stream.cuda_stream, cell_subdiv=args.subdiv, flags=flags,
                                  max_cells=1 << 30)
        self.sim.apply_params(self.p)
        self.sim.scene_block(*dims, SPACING, None, JITTER, SEED)
        self.sim.set_paused(False)
        st = self.sim.stats()
        self.grid_cells, self.sort_passes, self.subdiv = st.num_cells, st.sort_passes, st.cell_subdiv
        # hist + P onesweep + reorder + fill + rho + fused force/integrate + their two deferred-queue kernels
        self.launches_per_step = 7 + st.sort_passes
        self.fused = not (flags & sph.FLAG_NO_FUSE)
        self.parallelism = "1 process, 1 GPU"
        self.api = ("nprsph_upload_state (pos+vel, 32 B/particle) + nprsph_step(1) + nprsph_download_positions "
                    "(16 B/particle, asynchronous), pinned host buffers")

    def step(self, k):
        self.sim.step(k)

    def profile(self, k):
        return self.sim.profile_step(k)

    def nan_particles(self):
        return int(self.sim.stats().nan_particles)

    def close(self):
        self.sim.close()


class SlabRunner:
    """N > 1: ONE global dam break split into x slabs, one rank per GPU, ghost halo exchange +
    migration over NCCL send/recv every step, slab faces re-balanced as the fluid moves."""

    def __init__(self, sph, dims, args, rank, local, world, stream, torch, dist):
        from nprsph_b200.dist import SlabGroup, unique_id
        self.dims, self.world = dims, world
        self.n_total = dims[0] * dims[1] * dims[2]
        self.p = sph.scenes.dam_break_params(*dims)
        idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        self.grp = SlabGroup.nccl(rank, world, bytes(idt.cpu().numpy().tobytes()), device=local,
                                  stream=stream.cuda_stream, cell_subdiv=args.subdiv, max_cells=1 << 30,
                                  rebalance_every=-args.rebalance_every if args.balance == "time" else args.rebalance_every)
        self.grp.apply_params(self.p)
        self.grp.scene_block(*dims, SPACING, None, JITTER, SEED)
        self.grp.set_paused(False)
        self.grp.step(1)                       # distributes the scene; part of the warm-up
        info = self.grp.info()
        self.n_own = int(info.num_own)
        self.cap = int(info.cap_own)
        st = self.grp.sims[0].stats()
        self.subdiv = st.cell_subdiv
        self.grid_cells = (info.x_end - info.x_begin + 2 * st.cell_subdiv) * st.grid_dim[1] * st.grid_dim[2]
        self.sort_passes = int(info.sort_passes)                 # over the occupied x layers only
        # hist + P onesweep + gather/cells + 2 ghost cells + fill + rho (+deferred) + 3 x (fused force/integrate
        # + deferred) + counter set
        self.launches_per_step = 14 + self.sort_passes
        self.fused = True
        self.parallelism = (f"{world} slabs along x, 1 process/GPU, ghost halo (pos; v,rho) + migration "
                            f"via ncclSend/ncclRecv each step, faces re-balanced every {args.rebalance_every} steps "
                            f"by {'density-pass time' if args.balance == 'time' else 'particle count'}")
        self.api = ("nprsph_dist_upload_state (pos+vel, 32 B/particle) + nprsph_dist_step(1) + "
                    "nprsph_dist_download_positions (16 B/particle), pinned host buffers, every rank its own particles")

    def step(self, k):
        self.grp.step(k)

    def profile(self, k):
        return self.grp.profile_step(k)

    def nan_particles(self):
        return int(self.grp.info().nan_particles)

    def close(self):
        self.grp.close()


def timed(run, stream, torch, steps, warm, barrier, max_over_ranks, sampler=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run.step(max(warm, 3))
    barrier()
    if sampler:
        sampler.mark()
    e0.record(stream)
    run.step(steps)
    e1.record(stream)
    barrier()
    if sampler:
        sampler.mark()
    return max_over_ranks(e0.elapsed_time(e1)) / steps


def dist_parity(sph, rank, local, world, stream, torch, dist):
    """The slab-decomposed step over real NCCL against a single-context run of the same scene: a
    flow that pushes particles through every slab face (migration, ghosts, re-balancing)."""
    import numpy as np
    from nprsph_b200.dist import SlabGroup, unique_id
    nx, ny, nz, steps = 96, 32, 24, 100
    p = sph.scenes.dam_break_params(nx, ny, nz)
    p.gravity[0] = 300.0
    idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    grp = SlabGroup.nccl(rank, world, bytes(idt.cpu().numpy().tobytes()), device=local,
                         stream=stream.cuda_stream, cell_subdiv=2, rebalance_every=4)
    grp.apply_params(p)
    grp.scene_block(nx, ny, nz, SPACING, None, 2e-4, 21)
    grp.set_paused(False)
    grp.step(steps)
    rec, ids = grp.download()
    info = grp.info()
    mine = {"rec": rec, "ids": ids, "migrated": int(info.migrated_total), "rebalanced": int(info.rebalanced),
            "ghosts": int(info.ghosts_left + info.ghosts_right)}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    grp.close()
    if rank != 0:
        return None
    ref = sph.Simulation(device=local, stream=stream.cuda_stream, cell_subdiv=2)
    ref.apply_params(p)
    ref.scene_block(nx, ny, nz, SPACING, None, 2e-4, 21)
    ref.set_paused(False)
    ref.step(steps)
    want = ref.download()
    ref.close()
    n = nx * ny * nz
    got = np.full((n, 16), np.nan, np.float32)
    seen = np.zeros(n, np.int32)
    for g in gathered:
        seen[g["ids"]] += 1
        got[g["ids"]] = g["rec"]
    partition = bool((seen == 1).all())
    out = {"scene": f"{nx}x{ny}x{nz} block, g_x = 300, {steps} steps, NCCL send/recv, re-balancing every 4 steps "
                    f"vs nprsph_step on one context", "ranks": world, "ids_partition_exactly": partition,
           "migrated": sum(g["migrated"] for g in gathered), "ghosts": sum(g["ghosts"] for g in gathered),
           "face_moves": sum(g["rebalanced"] for g in gathered), "tolerance": 1e-5, "max_rel": {}}
    ok = partition and out["migrated"] > 0 and out["ghosts"] > 0
    for name, cols in (("pos", slice(0, 3)), ("vel", slice(4, 7)), ("force", slice(8, 11)), ("rho", slice(12, 13))):
        a, b = got[:, cols].astype(np.float64), want[:, cols].astype(np.float64)
        same_nan = bool(np.array_equal(np.isnan(a), np.isnan(b)))
        m = ~np.isnan(b)
        rel = float(np.abs(a[m] - b[m]).max() / max(np.abs(b[m]).max(), 1e-300)) if same_nan and m.any() else float("inf")
        out["max_rel"][name] = rel
        ok = ok and same_nan and rel <= 1e-5
    out["ok"] = bool(ok)
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import nprsph_b200 as sph

    rank, local, world = dist_env()
    if args.gpus > 1 and world == 1:
        relaunch_under_torchrun(args)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libnprsph has no CPU path")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    stream = torch.cuda.Stream(device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(ms):
        return reduce(ms, dist.ReduceOp.MAX) if world > 1 else float(ms)

    def sum_over_ranks(v):
        return reduce(v, dist.ReduceOp.SUM) if world > 1 else float(v)

    def make_runner(dims):
        if world == 1:
            return SingleRunner(sph, dims, args, local, stream)
        return SlabRunner(sph, dims, args, rank, local, world, stream, torch, dist)

    def slab_report(run, steps_timed, moved_before):
        """own min/max over ranks, hand-overs per step in the timed window, face moves so far"""
        if world == 1:
            return None
        info = run.grp.info()
        own = int(info.num_own)
        return {"own_min": int(reduce(own, dist.ReduceOp.MIN)), "own_max": int(reduce(own, dist.ReduceOp.MAX)),
                "migrated_per_step": (sum_over_ranks(int(info.migrated_total)) - moved_before) / max(steps_timed, 1),
                "face_moves_total": int(sum_over_ranks(int(info.rebalanced))),
                "nan_particles": int(sum_over_ranks(int(info.nan_particles)))}

    def moved_so_far(run):
        return sum_over_ranks(int(run.grp.info().migrated_total)) if world > 1 else 0.0

    parity = None
    if world > 1:
        parity = dist_parity(sph, rank, local, world, stream, torch, dist)

    # ---- main line ---------------------------------------------------------------------------
    side = args.side
    if world > 1 and args.scaling == "strong":
        dims, scaling = (512, 256, 512), "strong"                      # BASELINE configs[3]
    else:
        dims, scaling = (side * world, side, side), "weak"
    run = make_runner(dims)
    n_total = run.n_total
    warm = max(args.warmup, 3)
    peak, peak_src = peaks()

    lattice = None
    if args.evolve_steps > 0:          # the standing lattice first (secondary figure), then let the dam break
        ms_l = timed(run, stream, torch, min(args.steps, 20), warm, barrier, max_over_ranks)
        prof_l = run.profile(3)
        lattice = {"ms_per_step": round(ms_l, 4), "value": n_total / (ms_l * 1e-3), "unit": UNIT,
                   "per_kernel_ms": {k: round(v, 4) for k, v in prof_l.items()},
                   "step_frac": round(192 * n_total / world / (ms_l * 1e-3) / 1e9 / peak, 4),
                   "note": "the same scene before the dam breaks: every warp walks identical columns (the "
                           "friendliest arrangement; round 1's headline)"}
        run.step(args.evolve_steps)
    sampler = ClockSampler(local)
    sampler.start()                    # nvidia-smi needs ~0.1 s to deliver its first sample
    moved0 = moved_so_far(run)
    ms_per_step = timed(run, stream, torch, args.steps, warm, barrier, max_over_ranks, sampler)
    clocks = sampler.stop()
    value = n_total / (ms_per_step * 1e-3)
    slab = slab_report(run, args.steps + warm, moved0)
    nan_main = run.nan_particles() if world == 1 else slab["nan_particles"]

    # per-kernel device times (CUDA events between stages on the same stream), live
    prof = run.profile(max(3, min(args.steps, 10)))
    n_k = run.n_own if world == 1 else int(run.grp.info().num_own)
    per_rank = None
    if world > 1:                                   # every rank's stage times and slab size (skew)
        mine = {"rank": rank, "own": n_k, "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons"),
                **{k: round(v, 3) for k, v in prof.items() if v > 0}}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered
    # nprsph_step runs force + integrate as ONE launch (booked under "force")
    algo = dict(ALGO_BYTES_FUSED if run.fused else ALGO_BYTES)
    kernel_of = dict(KERNEL_OF_STAGE)
    if run.fused:
        kernel_of["force"] = "k_force_records<FUSE> (force + integrate passes in one launch)"
    walk = passes = None
    if world == 1:
        walk = run.sim.walk_stats()
        # per-pass times of the three-launch form on a second context carrying the same evolved state
        host = np.empty((run.n_own, 16), np.float32)
        run.sim.download(host)
        nofuse = SingleRunner(sph, dims, args, local, stream, flags=sph.FLAG_NO_FUSE)
        nofuse.sim.upload(host)
        nofuse.sim.step(3)
        passes = nofuse.profile(max(3, min(args.steps, 10)))
        nofuse.close()
        del host
    stage = max(algo, key=lambda k: prof[k])
    achieved = algo[stage] * n_k / (prof[stage] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{kernel_of[stage].split()[0]}@{n_k}@subdiv{run.subdiv}")
    roofline = {"kernel": kernel_of[stage], "bound": "hbm", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": algo[stage],
                "particles_per_launch": n_k, "kernel_ms": round(prof[stage], 4),
                "per_kernel_ms": {k: round(v, 4) for k, v in prof.items()},
                "per_kernel_frac": {k: round(algo[k] * n_k / (prof[k] * 1e-3) / 1e9 / peak, 4)
                                    for k in algo},
                "step_frac": round(192 * n_total / world / (ms_per_step * 1e-3) / 1e9 / peak, 4),
                "per_rank": per_rank,
                "three_launch_step": None if passes is None else {
                    "per_kernel_ms": {k: round(v, 4) for k, v in passes.items()},
                    "per_kernel_frac": {k: round(ALGO_BYTES[k] * n_k / (passes[k] * 1e-3) / 1e9 / peak, 4)
                                        for k in ALGO_BYTES},
                    "note": "same state, FLAG_NO_FUSE: k_rho, k_force_records, k_integrate as separate launches"},
                "note": "rho/force are instruction-issue bound (DESIGN.md 4); frac is algorithmic "
                        "bytes / time / measured HBM peak, rank 0"}
    if walk:
        # second figure of SURVEY 8(d): the neighbour passes in their own unit
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        t_rho = prof["rho"] * 1e-3
        tests_s = walk["distance_tests"] / t_rho
        issue_peak = 148 * 4 * sm_hz                          # warp instructions / s
        roofline["issue"] = {
            "distance_tests_per_step": walk["distance_tests"], "neighbours_per_particle": round(walk["neighbours"] / n_k, 2),
            "tests_per_particle": round(walk["distance_tests"] / n_k, 1), "columns_per_pair_walk": round(walk["columns"] / max(walk["pair_walks"] + walk["single_walks"], 1), 2),
            "unpaired_slots": walk["single_walks"],
            "k_rho_pair_tests_per_s": tests_s,
            "fp32_issue_peak_warp_inst_per_s": issue_peak,
            "k_rho_frac_of_issue_peak_spent_on_tests": round(tests_s * RHO_INSTR_PER_TEST / 32 / issue_peak, 4),
            "note": f"{RHO_INSTR_PER_TEST:.0f} issue slots per distance test (SASS count), 148 SMs x 4 schedulers x SM clock"}

    # ---- end to end through the C ABI with HOST buffers, every step -----------------------------
    if world == 1:
        n = run.n_own
        host = torch.empty(n * 16, dtype=torch.float32).pin_memory()
        run.sim.download_ptr(host.data_ptr(), n)
        rec = host.view(n, 16)
        hin = [(torch.empty(n * 4, dtype=torch.float32).pin_memory(), torch.empty(n * 4, dtype=torch.float32).pin_memory())
               for _ in range(2)]
        for hp, hv in hin:
            hp.view(n, 4).copy_(rec[:, 0:4]); hv.view(n, 4).copy_(rec[:, 4:8])
        hout = [torch.empty(n * 4, dtype=torch.float32).pin_memory() for _ in range(2)]

        def e2e_step(i):
            hp, hv = hin[i % 2]
            run.sim.upload_state_ptr(hp.data_ptr(), hv.data_ptr(), n)
            run.sim.step(1)
            run.sim.download_positions_ptr(hout[i % 2].data_ptr(), n, asynchronous=True)
            return n * 32, n * 16
        e2e_step(0)
        run.sim.sync()
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        for i in range(1, args.e2e_steps + 1):
            a, b = e2e_step(i)
            h2d, d2h = h2d + a, d2h + b
        run.sim.sync()                      # the last asynchronous download has landed
        e2e_ms = (time.perf_counter() - t0) * 1e3
        moved = float(np.nanmax(np.abs(hout[args.e2e_steps % 2].view(n, 4)[:, :3].numpy() - rec[:, 0:3].numpy())))
        assert 0.0 < moved < 0.1, f"e2e: downloaded positions are not one step away from the uploaded ones ({moved})"
        e2e_note = ("host wall clock around upload_state -> step -> asynchronous download_positions of every step, "
                    "final sync inside; the result of step i leaves while the inputs of step i+1 arrive")
        # PCIe yard-stick (plain torch copies of the same pinned buffers, both directions at once): the
        # floor of a step that moves these bytes, whatever the kernels cost
        d_in = torch.empty(n * 8, dtype=torch.float32, device="cuda")
        d_out = torch.empty(n * 4, dtype=torch.float32, device="cuda")
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        for it in range(2):                 # the first round warms the copy engines up
            p0.record()
            s_up.wait_event(p0); s_dn.wait_event(p0)
            for k in range(4):
                with torch.cuda.stream(s_up):
                    d_in[:n * 4].copy_(hin[k % 2][0], non_blocking=True); d_in[n * 4:].copy_(hin[k % 2][1], non_blocking=True)
                with torch.cuda.stream(s_dn):
                    hout[k % 2].copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_up); torch.cuda.current_stream().wait_stream(s_dn)
            p1.record(); torch.cuda.synchronize()
        pcie_ms = p0.elapsed_time(p1) / 4
        pcie = {"floor_ms_per_step": round(pcie_ms, 3), "h2d_GBps_with_d2h_running": round(n * 32 / pcie_ms / 1e6, 1),
                "what": "torch copies of the same pinned buffers (32 B/particle in, 16 B/particle out) on two streams, no kernels"}
        del d_in, d_out
    else:
        # slab mode: the same streaming pattern through nprsph_dist_upload_state / _download_positions
        # (pos4 = x, y, z, id bits; 32 B in and 16 B out per particle); every step starts from the same
        # uploaded state, as at N = 1
        cap = run.cap
        run.grp.freeze_faces(True)          # the same lists are uploaded every step: the faces must stay put
        rec0 = torch.empty(cap * 16, dtype=torch.float32).pin_memory()
        ids0 = torch.empty(cap, dtype=torch.int32).pin_memory()
        n0 = run.grp.download_ptr(0, rec0.data_ptr(), ids0.data_ptr(), cap)
        hp = torch.empty(n0 * 4, dtype=torch.float32).pin_memory()
        hv = torch.empty(n0 * 4, dtype=torch.float32).pin_memory()
        r0 = rec0.view(cap, 16)[:n0]
        hp.view(n0, 4).copy_(r0[:, 0:4]); hp.view(n0, 4)[:, 3] = ids0[:n0].view(torch.float32)
        hv.view(n0, 4).copy_(r0[:, 4:8])
        del rec0, ids0, r0
        hout = [torch.empty(cap * 4, dtype=torch.float32).pin_memory() for _ in range(2)]

        def e2e_step(i):
            run.grp.upload_state_ptr(0, hp.data_ptr(), hv.data_ptr(), n0)
            run.grp.step(1)
            # (synchronous: the overlapped form -- asynchronous=True, 2.67e9 /s at N = 2,
            #  profiles/bench_r2_n2_streaming_e2e.json -- was not re-measured at N = 8 after the face freeze)
            n_out = run.grp.download_positions_ptr(0, hout[i % 2].data_ptr(), cap, asynchronous=args.e2e_overlap)
            return n0 * 32, n_out * 16
        e2e_step(0)
        run.grp.sync()
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        for i in range(1, args.e2e_steps + 1):
            a, b = e2e_step(i)
            h2d, d2h = h2d + a, d2h + b
        run.grp.sync()                      # the last asynchronous download has landed
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        e2e_note = ("host wall clock, upload_state -> step -> download_positions of every rank's own particles "
                    f"({'asynchronous, final sync inside' if args.e2e_overlap else 'synchronous: copy in, step, copy out in series'}), "
                    "barrier on both sides")
    e2e_ms = max_over_ranks(e2e_ms)
    e2e = {"value": n_total * args.e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(sum_over_ranks(h2d) / args.e2e_steps),
           "d2h_bytes_per_step": int(sum_over_ranks(d2h) / args.e2e_steps),
           "steps": args.e2e_steps, "ms_per_step": round(e2e_ms / args.e2e_steps, 3), "api": run.api, "timing": e2e_note}
    if world == 1:
        e2e["pcie"] = pcie

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baselines(np.ascontiguousarray(rec.numpy()), run.p, run.subdiv)
    main_workload = workload_name(dims, args.evolve_steps, world)
    run.close()
    if world == 1:
        del host, hin, hout, rec
    else:
        del hp, hv, hout
    torch.cuda.empty_cache()

    # ---- BASELINE configs[3] (64 Mi strong scaling) and configs[4] (32 Mi per GPU weak scaling) ----
    configs = None
    if not args.no_extra_configs:
        configs = {}
        for name, cdims, kind in (("config4_strong_64Mi", (512, 256, 512), "strong"),
                                  ("config5_weak_32Mi_per_gpu", (256 * world, 256, 512), "weak")):
            if cdims == dims:
                configs[name] = {"same_as": "main line"}
                continue
            twin = next((k for k, v in configs.items() if v.get("dims") == list(cdims)), None)
            if twin:
                configs[name] = {"same_as": twin, "scaling": kind}
                continue
            r = make_runner(cdims)
            r.step(args.extra_evolve_steps)
            m0 = moved_so_far(r)
            k = max(5, min(args.steps, 20))
            ms = timed(r, stream, torch, k, 3, barrier, max_over_ranks)
            configs[name] = {"workload": workload_name(cdims, args.extra_evolve_steps, world), "scaling": kind, "dims": list(cdims),
                             "particles_total": r.n_total, "ms_per_step": round(ms, 4),
                             "value": r.n_total / (ms * 1e-3), "unit": UNIT, "steps": k,
                             "step_frac": round(192 * r.n_total / world / (ms * 1e-3) / 1e9 / peak, 4),
                             "slab": slab_report(r, k + 3, m0),
                             "nan_particles": r.nan_particles() if world == 1 else None}
            r.close()
            torch.cuda.empty_cache()

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": main_workload,
                          "particles_total": n_total, "cell_subdiv": int(run.subdiv),
                          "grid_cells_per_rank": int(run.grid_cells), "sort_passes": int(run.sort_passes),
                          "parallelism": run.parallelism, "nan_particles": nan_main,
                          "l2": "working set per step (>2 GB per GPU) exceeds the 126 MB L2; no flush needed"},
               "roofline": roofline, "e2e": e2e, "gpu_launches": run.launches_per_step * args.steps,
               "clocks": clocks}
        if cpu_baseline:
            out["cpu_baseline"] = cpu_baseline
        if lattice:
            out["lattice"] = lattice
        if slab:
            out["slab"] = slab
        if parity is not None:
            out["dist_parity"] = parity
        if configs:
            out["configs"] = configs
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ---- CPU legs (the only places bench.py touches oracle/) ------------------------------------------
def oracle_params(O, p):
    q = O.default_params()
    for k in ("mass", "smoothing_coeff", "visc", "resting_rho", "particle_radius", "gas_const", "damping", "dt", "pi"):
        setattr(q, k, getattr(p, k))
    for a in range(3):
        q.gravity[a] = p.gravity[a]
    for a in range(4):
        q.upper[a], q.lower[a] = p.upper[a], p.lower[a]
    return q


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def all_pairs_sample(O, host, q, budget_s):
    """Reference algorithm (all-pairs j loop, the shaders' own) for a bounded sample of the
    workload's particles on this box's host cores.  `host` already carries rho/p."""
    import numpy as np
    n = len(host)
    threads = O.num_threads()
    idx0 = np.linspace(0, n - 1, threads, dtype=np.int32)
    t0 = time.perf_counter(); O.sample_update(host, q, idx0); dt0 = time.perf_counter() - t0
    m = int(max(threads, min(4096, budget_s / max(dt0, 1e-6) * threads)))
    m -= m % threads
    idx = np.linspace(0, n - 1, m, dtype=np.int32)
    t0 = time.perf_counter(); O.sample_update(host, q, idx); dt = time.perf_counter() - t0
    return m / dt, m, dt


def egl_reference_run():
    """the reference's unmodified compute shaders run headless through EGL on this
    box's GPU (tools/egl_shader_runner.c), where the driver exposes GL and a checkout of the reference
    is named by NPRSPH_REFERENCE_SHADERS.  Reports why not otherwise."""
    import tempfile
    src = os.path.join(ROOT, "tools", "egl_shader_runner.c")
    shaders = os.environ.get("NPRSPH_REFERENCE_SHADERS", "")
    try:
        with tempfile.TemporaryDirectory() as tmp:
            exe = os.path.join(tmp, "egl_shader_runner")
            subprocess.check_call(["gcc", "-O2", "-std=c99", src, "-o", exe, "-ldl"], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
            r = subprocess.run([exe, shaders or os.path.join(tmp, "no-reference-checkout"), "200"], capture_output=True,
                               text=True, timeout=120)
    except (OSError, subprocess.SubprocessError) as e:
        return {"unavailable": f"runner did not build or run: {e}"}
    if r.returncode == 0:
        return json.loads(r.stdout.strip().splitlines()[-1])
    why = (r.stderr.strip().splitlines() or ["exit %d" % r.returncode])[-1]
    if r.returncode != 3 and not shaders:
        why = "EGL context available, but NPRSPH_REFERENCE_SHADERS does not name a checkout of the reference's NPR-SPH directory"
    return {"unavailable": why}


def cpu_baselines(host, p, subdiv):
    """Three CPU figures on this box's own cores: the reference's all-pairs algorithm on
    a bounded sample of the workload (`value`), the same UNIFORM-GRID algorithm as the GPU path on the
    whole workload (oracle_step_grid, bit-identical to all-pairs), and BASELINE configs[0] -- the
    reference's default 10,000-particle scene stepped by the OpenMP transcription of the shaders."""
    import numpy as np
    from oracle import oracle as O
    q = oracle_params(O, p)
    threads, n = O.num_threads(), len(host)
    v, m, dt = all_pairs_sample(O, host, q, budget_s=8.0)
    out = {"value": v, "unit": UNIT, "cores": threads, "cpu": cpu_model(), "kind": "port",
           "sample": f"oracle all-pairs update (rho+force+integrate, the reference's O(N) loop per "
                     f"particle) of {m} evenly spaced particles out of {n}, {dt:.1f} s"}
    G = host.copy()
    t0 = time.perf_counter(); O.step(G, q, 1, grid=int(subdiv)); dtg = time.perf_counter() - t0
    out["uniform_grid_same_algorithm"] = {
        "value": n / dtg, "unit": UNIT, "cores": threads, "steps": 1, "seconds": round(dtg, 2),
        "what": f"oracle_step_grid (cell = h/{subdiv}, the GPU path's neighbour search restated in C + OpenMP) over all {n} particles"}
    del G
    out["reference_shaders_egl"] = egl_reference_run()
    P = O.make_block(10, 100, 10)
    qd = O.default_params()
    O.step(P, qd, 3)
    k = 30
    t0 = time.perf_counter(); O.step(P, qd, k); dtc = time.perf_counter() - t0
    out["config1_default_scene_openmp"] = {
        "value": 10000 * k / dtc, "unit": UNIT, "cores": threads, "steps": k, "ms_per_step": round(dtc / k * 1e3, 3),
        "what": "BASELINE configs[0]: 10 x 100 x 10 block, the reference's constants, all-pairs OpenMP transcription (oracle_step)"}
    return out


def run_reference(args):
    """The reference's own algorithm on the host cores: the OpenMP C transcription of the three
    shaders (oracle/, all-pairs).  The GLSL + Win32 reference cannot be built here,
    so kind = "port".  Each step updates a bounded sample of the workload's particles."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 in its workers; the reference arm gets every host core
    if os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import numpy as np
    import nprsph_b200 as sph           # parameter recipe only (no CUDA call)
    from oracle import oracle as O
    side, gpus = args.side, max(args.gpus, 1)
    dims = (512, 256, 512) if gpus > 1 and args.scaling == "strong" else (side * gpus, side, side)
    n = dims[0] * dims[1] * dims[2]
    p = sph.scenes.dam_break_params(*dims)
    q = oracle_params(O, p)
    host = O.jitter(O.make_block(*dims), JITTER, SEED)
    host[:, 12] = p.resting_rho          # rho_j read by the force loop (timing only)
    threads = O.num_threads()
    steps, warm = args.steps, max(args.warmup, 1)
    idx0 = np.linspace(0, n - 1, threads, dtype=np.int32)
    t0 = time.perf_counter(); O.sample_update(host, q, idx0); dt0 = time.perf_counter() - t0
    per_step = min(6.0, 150.0 / (steps + warm))
    m = int(max(threads, per_step / max(dt0, 1e-6) * threads))
    m -= m % threads
    idx = np.linspace(0, n - 1, m, dtype=np.int32)
    for _ in range(warm):
        O.sample_update(host, q, idx)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.sample_update(host, q, idx)
    dt = time.perf_counter() - t0
    value = m * steps / dt
    sample = (f"{m} evenly spaced particles of the {n}-particle dam break per step, all-pairs "
              f"rho+force+integrate (oracle/sph_oracle.c), {threads} OpenMP threads on {cpu_model()}; the all-pairs "
              f"cost does not depend on the arrangement, so the block is not evolved first")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if dims == (512, 256, 512) and gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(dims, args.evolve_steps, gpus), "particles_total": n,
                   "algorithm": "all-pairs (reference)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "cpu": cpu_model(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        run_ours(a)
    else:
        # One rank failing must end the job, not hang it: the ordinary interpreter shutdown would wait in
        # nprsph_destroy for collectives whose peers never arrive, and torchrun only tears the other
        # ranks down once this process is gone.
        try:
            run_ours(a)
        except BaseException:
            import traceback
            traceback.print_exc()
            sys.stderr.flush(); sys.stdout.flush()
            os._exit(1)
