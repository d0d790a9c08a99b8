#!/usr/bin/env python
"""bench.py -- SPH particle-updates/s of libnprsph on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # reference algorithm on host cores

One "step" = one pass of the hot path (grid maintenance + rho/pressure + force + integrate,
i.e. display()'s compute block, Main.cpp:291-305) over every particle of the workload.
Workload: the 16,777,216-particle fp32 dam break (BASELINE.json configs[2], the configuration
the HBM-roofline headline is quoted on), 256^3 lattice block, stable parameter recipe of
SURVEY.md 8(d), seeded jitter.  Prints ONE JSON line (rank 0).
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
# algorithmic (compulsory) bytes per particle per launch, SURVEY.md 8(d) / DESIGN.md
ALGO_BYTES = {"rho": 32, "force": 64, "integrate": 96}
KERNEL_OF_STAGE = {"rho": "k_rho", "force": "k_force", "integrate": "k_integrate",
                   "sort": "k_onesweep(+k_radix_hist)", "reorder": "k_reorder_cells(+k_fill_gaps)",
                   "keys": "k_keys"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=256, help="lattice block is side^3 particles per GPU")
    ap.add_argument("--subdiv", type=int, default=2, help="grid cells per smoothing length")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--evolve-steps", type=int, default=2000,
                    help="N = 1: afterwards let the dam break run this many steps and time the step again on "
                         "the disordered fluid (reported under \"disordered\"; 0 = skip)")
    return ap.parse_args()


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
    """N = 1: one context, the plain C-ABI calls (nprsph_step / upload / download)."""

    def __init__(self, sph, O, args, local, stream):
        side = args.side
        self.n_own = side ** 3
        self.p = O.dam_break_params(side, side, side)
        self.sim = sph.Simulation(device=local, stream=stream.cuda_stream, cell_subdiv=args.subdiv)
        self.sim.apply_params(self.p)
        self.sim.scene_block(side, side, side, 0.005, None, 1e-4 * 0.005, 1234)
        self.sim.set_paused(False)
        st = self.sim.stats()
        self.grid_cells, self.sort_passes = st.num_cells, st.sort_passes
        # hist + P onesweep + reorder + fill + rho + fused force/integrate + their two deferred-queue kernels
        self.launches_per_step = 7 + st.sort_passes
        self.fused = True
        self._args, self._sph, self._local, self._stream = args, sph, local, stream
        self.parallelism = "1 process, 1 GPU"
        self.api = "nprsph_upload_particles + nprsph_step(1) + nprsph_download_particles, pinned host buffers"

    def step(self, k):
        self.sim.step(k)

    def profile(self, k):
        return self.sim.profile_step(k)

    def profile_passes(self, k):
        """Per-pass times of the three-launch form of the step (FLAG_NO_FUSE) on a second context
        carrying the same state, so that every pass keeps its own roofline line."""
        side = self._args.side
        sim = self._sph.Simulation(device=self._local, stream=self._stream.cuda_stream,
                                   cell_subdiv=self._args.subdiv, flags=self._sph.FLAG_NO_FUSE)
        sim.apply_params(self.p)
        sim.scene_block(side, side, side, 0.005, None, 1e-4 * 0.005, 1234)
        sim.set_paused(False)
        sim.step(5)
        prof = sim.profile_step(k)
        sim.close()
        return prof

    def alloc_host(self, torch):
        n = self.n_own
        self.h = [torch.empty(n * 16, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.sim.download_ptr(self.h[0].data_ptr(), n)

    def e2e_step(self, i):
        src, dst = self.h[i % 2], self.h[(i + 1) % 2]
        self.sim.upload_ptr(src.data_ptr(), self.n_own)
        self.sim.step(1)
        self.sim.download_ptr(dst.data_ptr(), self.n_own)
        return self.n_own * 64, self.n_own * 64, self.n_own

    def last_host_state(self, i):
        return self.h[i % 2].numpy().reshape(self.n_own, 16)


class SlabRunner:
    """N > 1: ONE global dam break (side*N x side x side particles) split into x slabs, one rank
    per GPU, ghost halo exchange + migration over NCCL send/recv every step."""

    def __init__(self, sph, O, args, rank, local, world, stream, torch, dist):
        from nprsph_b200.dist import SlabGroup, unique_id
        side = args.side
        self.world = world
        self.p = O.dam_break_params(side * world, side, side)
        idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        self.grp = SlabGroup.nccl(rank, world, bytes(idt.cpu().numpy().tobytes()), device=local,
                                  stream=stream.cuda_stream, cell_subdiv=args.subdiv)
        self.grp.apply_params(self.p)
        self.grp.scene_block(side * world, side, side, 0.005, None, 1e-4 * 0.005, 1234)
        self.grp.set_paused(False)
        self.grp.step(1)                       # distributes the scene; part of the warm-up
        info = self.grp.info()
        self.n_own = int(info.num_own)
        self.cap = int(info.cap_own)
        st = self.grp.sims[0].stats()
        self.grid_cells = (info.x_end - info.x_begin + 2 * st.cell_subdiv) * st.grid_dim[1] * st.grid_dim[2]
        self.sort_passes = int(info.sort_passes)                 # over the occupied x layers only
        # hist + P onesweep + gather/cells + 2 ghost cells + fill + rho + 3 force + integrate/classify
        # + the 4 deferred-queue kernels
        self.launches_per_step = 14 + self.sort_passes
        self.fused = False
        self.parallelism = (f"{world} slabs along x, 1 process/GPU, ghost halo (pos; v,rho) + migration "
                            f"via ncclSend/ncclRecv each step")
        self.api = "nprsph_dist_upload + nprsph_dist_step(1) + nprsph_dist_download, pinned host buffers"

    def step(self, k):
        self.grp.step(k)

    def profile(self, k):
        return self.grp.profile_step(k)

    def alloc_host(self, torch):
        self.h = [torch.empty(self.cap * 16, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.hid = [torch.empty(self.cap, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.n_host = self.grp.download_ptr(0, self.h[0].data_ptr(), self.hid[0].data_ptr(), self.cap)

    def e2e_step(self, i):
        a, b = i % 2, (i + 1) % 2
        n_in = self.n_host
        self.grp.upload_ptr(0, self.h[a].data_ptr(), self.hid[a].data_ptr(), n_in)
        self.grp.step(1)
        self.n_host = self.grp.download_ptr(0, self.h[b].data_ptr(), self.hid[b].data_ptr(), self.cap)
        return n_in * 68, self.n_host * 68, n_in

    def last_host_state(self, i):
        return self.h[i % 2].numpy().reshape(self.cap, 16)[:self.n_host]


def run_ours(args):
    import torch
    import torch.distributed as dist
    import nprsph_b200 as sph
    from oracle import oracle as O   # only for the parameter recipe and the cpu_baseline leg

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
    if world == 1:
        run = SingleRunner(sph, O, args, local, stream)
    else:
        run = SlabRunner(sph, O, args, rank, local, world, stream, torch, dist)
    n_total = args.side ** 3 * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    warm = max(args.warmup, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sampler.start()                    # nvidia-smi needs ~0.1 s to deliver its first sample
    run.step(warm)
    barrier()
    sampler.mark()
    e0.record(stream)
    run.step(args.steps)
    e1.record(stream)
    barrier()
    sampler.mark()
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = n_total * args.steps / (ms_total * 1e-3)

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
    # N = 1: nprsph_step runs force + integrate as ONE launch (booked under "force"); its
    # algorithmic bytes are those of the two passes it performs, 64 + 96.
    algo = dict(ALGO_BYTES)
    kernel_of = dict(KERNEL_OF_STAGE)
    passes = None
    if run.fused:
        algo = {"rho": 32, "force": 64 + 96}
        kernel_of["force"] = "k_force_records<FUSE> (force + integrate passes in one launch)"
        passes = run.profile_passes(max(3, min(args.steps, 10)))
    stage = max(algo, key=lambda k: prof[k])
    peak, peak_src = peaks()
    achieved = algo[stage] * n_k / (prof[stage] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{kernel_of[stage].split()[0]}@{n_k}@subdiv{args.subdiv}")
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

    # end to end through the C ABI with HOST buffers: upload -> step -> download, every step
    run.alloc_host(torch)
    run.e2e_step(0)
    barrier()
    e0.record(stream)
    h2d = d2h = 0
    for i in range(1, args.e2e_steps + 1):
        a, b, _ = run.e2e_step(i)
        h2d, d2h = h2d + a, d2h + b
    e1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": n_total * args.e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(sum_over_ranks(h2d) / args.e2e_steps),
           "d2h_bytes_per_step": int(sum_over_ranks(d2h) / args.e2e_steps),
           "steps": args.e2e_steps, "api": run.api}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import numpy as np
        host = np.ascontiguousarray(run.last_host_state(args.e2e_steps + 1))   # carries rho/p
        cpu_baseline = cpu_sample(O, host, run.p, budget_s=12.0)

    # The timed region above still sees the initial lattice (warm-up + steps are ~100 steps).  A real
    # fluid is disordered: the same scene a few thousand steps later, same metric, same kernels.
    disordered = None
    if world == 1 and args.evolve_steps > 0:
        run.step(args.evolve_steps)
        k = 20
        run.step(3)
        e0.record(stream)
        run.step(k)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k
        prof_d = run.profile(5)
        disordered = {"evolved_steps": args.evolve_steps + args.steps + warm, "ms_per_step": round(ms, 4),
                      "value": n_total / (ms * 1e-3), "unit": UNIT,
                      "per_kernel_ms": {k_: round(v, 4) for k_, v in prof_d.items()},
                      "nan_particles": int(run.sim.stats().nan_particles),
                      "note": "same scene and kernels after the dam has broken: lanes of a warp no longer "
                              "walk identical columns (DESIGN.md 4, Disordered arrangements)"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": f"dam_break_{args.side ** 3}_per_gpu_fp32 (block {args.side * world}x"
                                      f"{args.side}x{args.side}, h=2s, stable recipe SURVEY 8(d), jitter seed 1234)",
                          "particles_total": n_total, "cell_subdiv": args.subdiv,
                          "grid_cells_per_rank": int(run.grid_cells), "sort_passes": int(run.sort_passes),
                          "parallelism": run.parallelism,
                          "l2": "working set per step (>2 GB per GPU) exceeds the 126 MB L2; no flush needed"},
               "roofline": roofline, "e2e": e2e, "gpu_launches": run.launches_per_step * args.steps,
               "clocks": clocks}
        if cpu_baseline:
            out["cpu_baseline"] = cpu_baseline
        if disordered:
            out["disordered"] = disordered
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_sample(O, host, p, budget_s):
    """Reference algorithm (all-pairs j loop, the shaders' own) for a bounded sample of the
    workload's particles on this box's host cores.  `host` already carries rho/p."""
    import numpy as np
    n = len(host)
    threads = O.num_threads()
    idx0 = np.linspace(0, n - 1, threads, dtype=np.int32)
    t0 = time.perf_counter(); O.sample_update(host, p, idx0); dt0 = time.perf_counter() - t0
    m = int(max(threads, min(4096, budget_s / max(dt0, 1e-6) * threads)))
    m -= m % threads
    idx = np.linspace(0, n - 1, m, dtype=np.int32)
    t0 = time.perf_counter(); O.sample_update(host, p, idx); dt = time.perf_counter() - t0
    return {"value": m / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"oracle all-pairs update (rho+force+integrate, the reference's O(N) loop per "
                      f"particle) of {m} evenly spaced particles out of {n}, {dt:.1f} s"}


def run_reference(args):
    """The reference's own algorithm on the host cores: the OpenMP C transcription of the three
    shaders (oracle/, all-pairs).  The GLSL + Win32 reference cannot be built here,
    so kind = "port".  Each step updates a bounded sample of the workload's particles."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    side = args.side
    n = side ** 3
    p = O.dam_break_params(side, side, side)
    host = O.jitter(O.make_block(side, side, side), 1e-4 * 0.005, 1234)
    host[:, 12] = p.resting_rho          # rho_j read by the force loop (timing only)
    threads = O.num_threads()
    steps, warm = args.steps, max(args.warmup, 1)
    idx0 = np.linspace(0, n - 1, threads, dtype=np.int32)
    t0 = time.perf_counter(); O.sample_update(host, p, idx0); dt0 = time.perf_counter() - t0
    per_step = min(6.0, 150.0 / (steps + warm))
    m = int(max(threads, per_step / max(dt0, 1e-6) * threads))
    m -= m % threads
    idx = np.linspace(0, n - 1, m, dtype=np.int32)
    for _ in range(warm):
        O.sample_update(host, p, idx)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.sample_update(host, p, idx)
    dt = time.perf_counter() - t0
    value = m * steps / dt
    sample = (f"{m} evenly spaced particles of the {n}-particle dam break per step, all-pairs "
              f"rho+force+integrate (oracle/sph_oracle.c), {threads} OpenMP threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"dam_break_{n}_per_gpu_fp32 (side {side}, h=2s, stable recipe SURVEY 8(d), "
                               f"jitter seed 1234)", "particles_total": n, "algorithm": "all-pairs (reference)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
