import sys, json
import numpy as np
sys.path.insert(0, ".")
import nprsph_b200 as sph

side=128
p = sph.scenes.dam_break_params(side, side, side)
sim = sph.Simulation(cell_subdiv=2, flags=sph.FLAG_COUNT_NEIGHBOURS)
sim.apply_params(p)
sim.scene_block(side, side, side, 0.005, None, 1e-4 * 0.005, 1234)
sim.set_paused(False)
for cp in (5, 1000, 2000, 4000):
    sim.step(cp - sim.stats().steps_done)
    c = sim.debug_read(sph.DBG_COUNTS_RHO).astype(np.float64)
    G = sim.download()
    rho = G[:, 12]
    ok = ~np.isnan(G[:, 0])
    # occupancy of cells: particles per occupied cell
    keys = sim.debug_read(sph.DBG_SORTED_KEYS)
    u, cnt = np.unique(keys, return_counts=True)
    print(json.dumps({"steps": cp, "nbr_mean": c[ok].mean(), "nbr_p50": float(np.percentile(c[ok], 50)), "nbr_p99": float(np.percentile(c[ok], 99)), "nbr_max": c[ok].max(),
      "rho_mean": float(rho[ok].mean()), "rho_p99": float(np.percentile(rho[ok], 99)), "per_cell_mean": float(cnt.mean()), "per_cell_p99": float(np.percentile(cnt, 99)), "per_cell_max": int(cnt.max()),
      "speed_max": float(np.abs(G[ok, 4:7]).max())}))
