"""Scratch: evolve a dam break, then run a few more steps (ncu: skip the evolution launches)."""
import sys
sys.path.insert(0, ".")
import nprsph_b200 as sph
from oracle import oracle as O
side, steps = int(sys.argv[1]), int(sys.argv[2])
p = O.dam_break_params(side, side, side)
sim = sph.Simulation(cell_subdiv=2)
sim.apply_params(p)
sim.scene_block(side, side, side, 0.005, None, 1e-4 * 0.005, 1234)
sim.set_paused(False)
sim.step(steps); sim.sync()
print("evolved", sim.stats().nan_particles)
sim.step(2); sim.sync()
