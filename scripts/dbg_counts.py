import sys
import numpy as np
sys.path.insert(0, ".")
import nprsph_b200 as sph
from oracle import oracle as O
p = O.default_params()
P = O.make_block(10, 100, 10)
subdiv = 2
sim = sph.Simulation(cell_subdiv=subdiv, flags=sph.FLAG_COUNT_NEIGHBOURS)
sim.apply_params(p)
for it in range(3):
    for fresh in (0, 1):
        s = sim
        if fresh:
            s = sph.Simulation(cell_subdiv=subdiv, flags=sph.FLAG_COUNT_NEIGHBOURS); s.apply_params(p)
        s.upload(P)
        s.pass_rho()
        Q = P.copy()
        c = O.pass_rho(Q, p, counts=True)
        g = s.debug_read(sph.DBG_COUNTS_RHO)
        ids = s.debug_read(sph.DBG_SLOT_IDS)
        slot_of = np.empty_like(ids); slot_of[ids] = np.arange(len(ids), dtype=ids.dtype)
        bad = np.nonzero(g != c)[0]
        print("iter", it, "fresh", fresh, "mismatches", len(bad), "of", len(c))
        for i in bad[:12]:
            d = P[:, :3].astype(np.float32) - P[i, :3]
            r2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            near = np.argsort(np.abs(r2 - np.float32(0.0004)))[:2]
            print(" ", i, "slot", slot_of[i], "gpu", g[i], "oracle", c[i], "pos", P[i, :3], "closest-to-h2", near, r2[near], slot_of[near])
    O.step(P, p, 1)
