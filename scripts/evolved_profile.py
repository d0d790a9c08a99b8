"""Per-stage times of the step on an EVOLVED dam break (disordered fluid instead of the initial
lattice) and the share of slot pairs that share one walk (run under gpurun)."""
import os, sys, json
import numpy as np
sys.path.insert(0, ".")
import nprsph_b200 as sph
from oracle import oracle as O

side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
checkpoints = [int(a) for a in sys.argv[2:]] or [0, 500, 2000]
p = O.dam_break_params(side, side, side)
flags = int(os.environ.get("EV_FLAGS", "0"))
sim = sph.Simulation(cell_subdiv=2, flags=flags)
sim.apply_params(p)
sim.scene_block(side, side, side, 0.005, None, 1e-4 * 0.005, 1234)
sim.set_paused(False)
done = 0
for cp in checkpoints:
    sim.step(cp - done); done = cp
    sim.step(3)
    prof = sim.profile_step(5); done += 8
    n = sim.num_particles
    half = (n + 1) // 2
    ctl = np.empty(half, np.uint32)
    sim._ck(sim.lib.nprsph_debug_read(sim._h, 6, ctl.ctypes.data, ctl.nbytes))
    st = sim.stats()
    print(json.dumps({"flags": flags, "steps": done, "pair_walk_frac": float((ctl & 1).mean()), "rescan_frac": float(((ctl & 24) != 0).mean()),
                      "nan": st.nan_particles, "ms": {k: round(v, 3) for k, v in prof.items()},
                      "step_ms": round(sum(prof.values()), 3)}))
