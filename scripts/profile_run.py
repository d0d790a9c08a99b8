"""A dam break of side^3 particles: EVOLVE steps to get a real (disordered) fluid, then STEPS more
steps -- the launches ncu captures (skip the evolution with -s / --launch-skip).

    python scripts/profile_run.py SIDE EVOLVE STEPS [SUBDIV]
"""
import sys
sys.path.insert(0, ".")
import nprsph_b200 as sph

side, evolve, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
subdiv = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sim = sph.Simulation(cell_subdiv=subdiv)
sim.apply_params(sph.scenes.dam_break_params(side, side, side))
sim.scene_block(side, side, side, 0.005, None, 1e-4 * 0.005, 1234)
sim.set_paused(False)
sim.step(evolve); sim.sync()
print("evolved", evolve, "steps; nan particles:", sim.stats().nan_particles, flush=True)
sim.step(steps); sim.sync()
print("done", sim.num_particles)
