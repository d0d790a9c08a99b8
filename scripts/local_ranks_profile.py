"""Scratch: 4 virtual ranks on ONE GPU (LOCAL transport); per-launch k_rho/k_force times under ncu
show whether a slab's kernels are slower for geometric reasons."""
import sys
sys.path.insert(0, ".")
import nprsph_b200 as sph
from nprsph_b200.dist import SlabGroup

side, world = int(sys.argv[1]), int(sys.argv[2])
jit = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
subdiv = int(sys.argv[4]) if len(sys.argv) > 4 else 2
p = sph.scenes.dam_break_params(side * world, side, side)
g = SlabGroup.local(world, cell_subdiv=subdiv)
g.apply_params(p)
g.scene_block(side * world, side, side, 0.005, None, jit * 0.005, 1234)
g.set_paused(False)
g.step(4)
g.sync()
print([int(g.info(w).num_own) for w in range(world)])
