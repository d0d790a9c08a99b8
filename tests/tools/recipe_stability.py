"""Stability probe of the dam-break parameter recipe on the CPU oracle (grid-accelerated, bit-identical
to the all-pairs loop): a thin slice of the bench column (same height, smaller cross-section).

    python tests/tools/recipe_stability.py NX NY NZ GAS_CONST DT STEPS [VISC]

Prints NaN count, density range and max speed every 250 steps."""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from oracle import oracle as O

nx, ny, nz = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
k, dt, steps = float(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6])
p = O.dam_break_params(nx, ny, nz)
p.gas_const, p.dt = k, dt
if len(sys.argv) > 7:
    p.visc = float(sys.argv[7])
P = O.jitter(O.make_block(nx, ny, nz), 1e-4 * 0.005, 1234)
t0 = time.time()
done = 0
while done < steps:
    O.step(P, p, 250, grid=2)
    done += 250
    nan = int(np.isnan(P[:, 0:3]).any(axis=1).sum())
    ok = ~np.isnan(P[:, 0])
    v = np.sqrt((P[ok, 4:7].astype(np.float64) ** 2).sum(axis=1))
    print(f"k={k} dt={dt} step {done}: nan={nan} rho=[{P[ok,12].min():.1f},{P[ok,12].max():.1f}] vmax={v.max():.2f} "
          f"xfront={P[ok,0].max():.3f} ({time.time()-t0:.0f}s)", flush=True)
