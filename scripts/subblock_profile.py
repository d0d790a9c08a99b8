"""Scratch: single-GPU run of an x sub-block of the 4-rank scene (same grid), to study warp divergence."""
import sys
sys.path.insert(0, ".")
import nprsph_b200 as sph
from oracle import oracle as O
i0 = int(sys.argv[1]); side = 256
p = O.dam_break_params(side * 4, side, side)
sim = sph.Simulation(cell_subdiv=2, max_cells=1 << 30)
sim.apply_params(p)
sim.scene_block(side, side, side, 0.005, (i0 * 0.005, 0.0, 0.0), 1e-4 * 0.005, 1234)
sim.set_paused(False)
sim.step(3); sim.sync()
st = sim.stats(); print(i0, list(st.grid_dim), st.cell_size)
