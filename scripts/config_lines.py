"""Secondary measurement lines the judge asked for, one JSON line each:
  config1   BASELINE configs[0]: the reference's 10,000-particle default scene (h = 4 lattice spacings),
            automatic cell subdivision (4): ms per step, per-kernel times, walk statistics, the share
            of slots that replay records vs re-test candidates
  h4s_16M   16,777,216-particle dam break at the reference's smoothing ratio h = 4 s (~250
            neighbours): the compute-bound regime of SURVEY 8(d)
    python scripts/config_lines.py [config1|h4s_16M] ..."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
import torch
import nprsph_b200 as sph

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0


def timed(sim, steps):
    sim.step(5); sim.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(sim.stream)
    e0.record(st); sim.step(steps); e1.record(st); sim.sync()
    return e0.elapsed_time(e1) / steps


def line(name, sim, n, steps, extra):
    ms = timed(sim, steps)
    prof = sim.profile_step(5)
    ws = sim.walk_stats()
    st = sim.stats()
    out = {"config": name, "particles": n, "ms_per_step": round(ms, 4), "particle_updates_per_s": n / (ms * 1e-3),
           "cell_subdiv": int(st.cell_subdiv), "steps_done": int(st.steps_done), "nan_particles": int(st.nan_particles),
           "per_kernel_ms": {k: round(v, 4) for k, v in prof.items() if v > 0.0005},
           "step_frac_of_measured_hbm": round(192 * n / (ms * 1e-3) / 1e9 / PEAK, 5),
           "rho_frac": round(32 * n / (prof["rho"] * 1e-3) / 1e9 / PEAK, 5),
           "force_frac_fused_96B": round(96 * n / (prof["force"] * 1e-3) / 1e9 / PEAK, 5),
           "tests_per_particle": round(ws["distance_tests"] / n, 1), "neighbours_per_particle": round(ws["neighbours"] / n, 1),
           "pair_walk_share": round(2 * ws["pair_walks"] / n, 4), "k_rho_pair_tests_per_s": ws["distance_tests"] / (prof["rho"] * 1e-3)}
    out.update(extra)
    print(json.dumps(out), flush=True)


which = sys.argv[1:] or ["config1", "h4s_16M"]
if "config1" in which:
    sim = sph.Simulation()                      # the reference's scene and constants are the defaults
    sim.set_paused(False)
    line("config1_default_scene_fresh", sim, 10000, 20, {"note": "steps 5-25 of the reference scene (an explosion: rho = 160 rho0)"})
    sim.step(300)
    line("config1_default_scene_step300", sim, 10000, 50, {"note": "gas-like steady state after 300 steps (NaN particles as in the reference, SURVEY App. C)"})
    sim.close()
if "h4s_16M" in which:
    side = 256
    p = sph.scenes.dam_break_params(side, side, side)
    p.smoothing_coeff = 4.0
    p.mass = 1.25e-4                            # lattice-consistent mass at h = 4 s
    sim = sph.Simulation()
    sim.apply_params(p)
    sim.scene_block(side, side, side, 0.005, None, 1e-4 * 0.005, 1234)
    sim.set_paused(False)
    line("dam_break_16M_h4s_lattice", sim, side ** 3, 5, {"note": "h = 4 lattice spacings (the reference's ratio), cell = h/4, 81 columns per walk"})
    sim.step(300)
    line("dam_break_16M_h4s_step300", sim, side ** 3, 5, {})
    sim.close()
