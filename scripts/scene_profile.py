"""Per-kernel times, walk statistics and density range of a dam break of NX x NY x NZ particles at a
list of step counts.   python scripts/scene_profile.py NX NY NZ STEP [STEP ...]"""
import json, sys
import numpy as np
sys.path.insert(0, ".")
import nprsph_b200 as sph

nx, ny, nz = (int(a) for a in sys.argv[1:4])
marks = [int(a) for a in sys.argv[4:]]
sim = sph.Simulation(max_cells=1 << 30)
sim.apply_params(sph.scenes.dam_break_params(nx, ny, nz))
sim.scene_block(nx, ny, nz, 0.005, None, 1e-4 * 0.005, 1234)
sim.set_paused(False)
n = nx * ny * nz
for m in marks:
    sim.step(max(m - int(sim.stats().steps_done), 0))
    prof = sim.profile_step(5)
    ws = sim.walk_stats()
    st = sim.stats()
    P = sim.download()
    ok = ~np.isnan(P[:, 0])
    v = np.sqrt((P[ok, 4:7].astype(np.float64) ** 2).sum(axis=1))
    print(json.dumps({"dims": [nx, ny, nz], "steps": int(st.steps_done), "grid": list(st.grid_dim), "sort_passes": int(st.sort_passes),
                      "ms": {k: round(x, 3) for k, x in prof.items() if x > 0.001}, "step_ms": round(sum(prof.values()), 3),
                      "tests_per_particle": round(ws["distance_tests"] / n, 1), "neighbours": round(ws["neighbours"] / n, 2),
                      "unpaired": ws["single_walks"], "nan": int(st.nan_particles),
                      "rho_max": float(P[ok, 12].max()), "rho_p999": float(np.percentile(P[ok, 12], 99.9)),
                      "v_max": float(v.max()), "x_front": float(P[ok, 0].max())}), flush=True)
    del P
