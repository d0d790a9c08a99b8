"""How much work the neighbour walks of an arrangement hold, and how well a warp's lanes agree:
candidates per pair walk (5x5 columns, +-2 cells in z around the pair, no culling) and the
lock-step cost sum_c max_lane ceil(len_c / 2) against the mean lane's sum_c ceil(len_c / 2).
Lattice vs evolved dam break (run under gpurun; the analysis itself is numpy)."""
import sys, json
import numpy as np
sys.path.insert(0, ".")
import nprsph_b200 as sph


side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = sph.scenes.dam_break_params(side, side, side)
sim = sph.Simulation(cell_subdiv=2)
sim.apply_params(p)
sim.scene_block(side, side, side, 0.005, None, 1e-4 * 0.005, 1234)
sim.set_paused(False)
rng = np.random.default_rng(0)
for steps in (4, 4000):
    sim.step(steps - sim.stats().steps_done)
    st = sim.stats()
    dx, dy, dz = [int(v) for v in st.grid_dim]
    keys = sim.debug_read(sph.DBG_SORTED_KEYS).astype(np.int64)
    cs = sim.debug_read(sph.DBG_CELL_START).astype(np.int64)
    n = len(keys)
    valid = keys < dx * dy * dz
    nv = int(valid.sum())
    warps = rng.choice(nv // 64 - 1, size=min(3000, nv // 64 - 1), replace=False)
    ratios, cands, pairable = [], [], []
    for w in warps:
        a = keys[w * 64: w * 64 + 64: 2]; b = keys[w * 64 + 1: w * 64 + 64: 2]
        ax, ay, az = a // (dy * dz), (a // dz) % dy, a % dz
        bx, by, bz = b // (dy * dz), (b // dz) % dy, b % dz
        ok = (ax == bx) & (ay == by) & (np.abs(az - bz) <= 1)
        pairable.append(ok.mean())
        zlo, zhi = np.minimum(az, bz) - 2, np.maximum(az, bz) + 2
        lens = np.zeros((32, 25), np.int64)
        c = 0
        for ox in range(-2, 3):
            for oy in range(-2, 3):
                x, y = ax + ox, ay + oy
                inb = (x >= 0) & (x < dx) & (y >= 0) & (y < dy)
                row = (np.clip(x, 0, dx - 1) * dy + np.clip(y, 0, dy - 1)) * dz
                l = cs[row + np.clip(zhi, 0, dz - 1) + 1] - cs[row + np.clip(zlo, 0, dz - 1)]
                lens[:, c] = np.where(inb & ok, l, 0)
                c += 1
        trips = (lens + 1) // 2
        lock = trips.max(axis=0).sum()
        mean = trips.sum(axis=1)[ok].mean() if ok.any() else 1
        ratios.append(lock / max(mean, 1))
        cands.append(lens.sum(axis=1)[ok].mean() if ok.any() else 0)
    print(json.dumps({"steps": steps, "pairable": float(np.mean(pairable)), "candidates_per_pair_walk_unculled": float(np.mean(cands)),
                      "lockstep_over_mean_lane": float(np.mean(ratios)), "p90": float(np.percentile(ratios, 90))}))
