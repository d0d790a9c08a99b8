"""Scratch: a few steps at a given size for ncu captures (run under gpurun + ncu)."""
import sys
sys.path.insert(0, ".")
import nprsph_b200 as sph
from oracle import oracle as O
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
subdiv = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
p = O.dam_break_params(n1, n1, n1)
sim = sph.Simulation(cell_subdiv=subdiv)
sim.apply_params(p)
sim.scene_block(n1, n1, n1, 0.005, None, 1e-4 * 0.005, 1234)
sim.set_paused(False)
sim.step(steps)
sim.sync()
print("done", sim.num_particles)
