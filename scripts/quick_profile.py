"""Scratch: per-stage timings of the step at a given size (run under gpurun)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import nprsph_b200 as sph
from oracle import oracle as O

def run(n1, subdiv, steps=10, coeff=2.0, flags=0):
    nx = ny = nz = n1
    p = O.dam_break_params(nx, ny, nz)
    p.smoothing_coeff = coeff
    sim = sph.Simulation(cell_subdiv=subdiv, flags=flags)
    sim.apply_params(p)
    sim.scene_block(nx, ny, nz, 0.005, None, 1e-4 * 0.005, 1234)
    sim.set_paused(False)
    sim.step(5); sim.sync()
    t = time.time(); sim.step(steps); sim.sync(); dt = (time.time() - t) / steps
    prof = sim.profile_step(steps)
    st = sim.stats()
    n = sim.num_particles
    print(json.dumps({"n": n, "subdiv": subdiv, "flags": flags, "coeff": coeff, "ms_per_step_wall": dt * 1e3,
                      "updates_per_s": n / dt, "stages_ms": prof, "cells": st.num_cells,
                      "dim": list(st.grid_dim), "key_bits": st.key_bits, "nan": st.nan_particles}))

if __name__ == "__main__":
    import os
    only = os.environ.get("QP_ONLY")            # "subdiv,flags" to run a single configuration
    for n1 in [int(a) for a in sys.argv[1:]] or [100, 256]:
        for subdiv in (1, 2):
            for flags in [int(f) for f in os.environ.get("QP_FLAGS", "0,2").split(",")]:
                if only and only != f"{subdiv},{flags}":
                    continue
                run(n1, subdiv, flags=flags)
