"""Where and when the dam-break scene grows NaN particles (two particles clamped onto the same wall
corner -> normalize(0)): NaN count every CHUNK steps, lattice indices of the
NaN particles.   python scripts/nan_probe.py SIDE STEPS CHUNK [wall_gap_in_spacings] [gas_const]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import nprsph_b200 as sph

side, steps, chunk = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
gap = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
p = sph.scenes.dam_break_params(side, side, side)
if len(sys.argv) > 5:
    p.gas_const = float(sys.argv[5])
s = 0.005
p.lower = [-gap * s, -gap * s, -gap * s, 1.0]
p.upper[2] = float(np.float32(side * s + (gap - 1.0) * s + s))        # symmetric gap behind the last lattice plane
sim = sph.Simulation(cell_subdiv=2)
sim.apply_params(p)
sim.scene_block(side, side, side, s, None, 1e-4 * s, 1234)
sim.set_paused(False)
done, seen = 0, set()
pos = np.empty((side ** 3, 4), np.float32)
while done < steps:
    sim.step(chunk); done += chunk
    n = int(sim.stats().nan_particles)
    line = f"gap {gap} k {p.gas_const}: step {done}: nan {n}"
    if n != len(seen):
        sim.download_positions(pos)
        ids = np.flatnonzero(np.isnan(pos[:, 0]))
        new = [int(i) for i in ids if int(i) not in seen]
        seen.update(new)
        line += " new (i,j,k): " + " ".join(str((i // (side * side), (i // side) % side, i % side)) for i in new[:12])
    print(line, flush=True)
