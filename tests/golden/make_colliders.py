"""Regression fixture for the collider extension: the reference has no
obstacles (README.md:59 lists them as future work), so there is nothing of the reference to pin
against -- this freezes the repo's own specification (oracle/sph_oracle.c:collide_one) so that
the oracle and the CUDA path cannot drift together unnoticed.

    python tests/golden/make_colliders.py        # rewrites tests/golden/colliders.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def scene():
    nx, s = 10, 0.005
    p = O.dam_break_params(nx, nx, nx)
    p.gas_const = 200.0            # the fixture was frozen with the softer stiffness of SURVEY 8(d)
    cols = [("sphere", (0.07, 0.015, 0.03), 0.012), ("box", (0.05, -s, 0.0), (0.056, 0.02, 0.02))]
    P = O.jitter(O.make_block(nx, nx, nx), 0.2 * s, seed=21)
    P[:, 4] = 1.5
    return p, cols, P


def run(p, cols, P, steps):
    oc = [O.sphere(c[1], c[2]) if c[0] == "sphere" else O.box(c[1], c[2]) for c in cols]
    out = []
    for s in range(steps):
        O.pass_rho(P, p)
        O.pass_force(P, p)
        O.pass_integrate(P, p, oc)
        if (s + 1) % 20 == 0:
            out.append(P.copy())
    return np.stack(out)


if __name__ == "__main__":
    p, cols, P = scene()
    states = run(p, cols, P.copy(), 120)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "colliders.npz"), start=P, states=states)
    d = np.linalg.norm(states[-1][:, :3] - np.array(cols[0][1]), axis=1)
    print("states", states.shape, "closest approach to the sphere centre / R:", d.min() / cols[0][2])
