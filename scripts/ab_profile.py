"""Per-kernel times of the step for the library in NPRSPH_LIB (A/B builds), 16 Mi dam break:
standing lattice and after EVOLVE steps.   python scripts/ab_profile.py [SIDE] [EVOLVE] [TAG]"""
import json, os, sys
sys.path.insert(0, ".")
import nprsph_b200 as sph

side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
evolve = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
tag = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(os.environ.get("NPRSPH_LIB", "default"))
sim = sph.Simulation()
sim.apply_params(sph.scenes.dam_break_params(side, side, side))
sim.scene_block(side, side, side, 0.005, None, 1e-4 * 0.005, 1234)
sim.set_paused(False)
out = {"lib": tag, "side": side}
for name, steps in (("lattice", 5), (f"evolved{evolve}", evolve)):
    sim.step(steps)
    sim.profile_step(3)
    prof = sim.profile_step(10)
    out[name] = {k: round(v, 4) for k, v in prof.items() if v > 0.001}
    out[name]["step"] = round(sum(prof.values()), 4)
out["nan"] = int(sim.stats().nan_particles)
print(json.dumps(out), flush=True)
