"""Scene and parameter recipes of the workloads BASELINE.json names (host-side mirror of the
reference's constants, Main.cpp:33-36,110-122 and the shader consts).  Plain data: the field names
are the reference's, so `Simulation.apply_params()` takes a record as it is.  Nothing here computes
physics; the oracle has its own copy of the defaults (oracle/sph_oracle.c) and a test holds the two
equal."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Params:
    # ConstantsUniform (Main.cpp:110-116)
    mass: float = 0.02
    smoothing_coeff: float = 4.0
    visc: float = 3000.0
    resting_rho: float = 1000.0
    # BoundaryUniform (Main.cpp:118-122)
    upper: list = field(default_factory=lambda: [0.5, 1.0, 0.5, 1.0])
    lower: list = field(default_factory=lambda: [-0.1, -0.35, -0.1, 1.0])
    # compile-time constants of the reference (Main.cpp:33-36; rho_pres_comp.glsl:5,8,33;
    # force_comp.glsl:33; integrate_comp.glsl:8,33)
    particle_radius: float = 0.005
    gas_const: float = 2000.0
    gravity: list = field(default_factory=lambda: [0.0, -9806.65, 0.0])
    damping: float = 0.3
    dt: float = 1.0e-4
    pi: float = 3.141592741


def default_params() -> Params:
    """The reference's own scene: 10 x 100 x 10 block, h = 4 lattice spacings (BASELINE configs[0])."""
    return Params()


DEFAULT_BLOCK = (10, 100, 10)

# Dam-break recipe (BASELINE configs[1..4]).  SURVEY.md 8(d) validated mass / viscosity / gravity for
# h = 2 lattice spacings with a stiffness of 200 (sound speed 14 m/s).  That is too soft for the 1.28 m
# column of the 16 Mi scene: the collapse reaches 4-5 m/s, i.e. Mach 0.3 -- densities of 1.45 rho0 and
# particles pressed onto the same wall corner (normalize(0) = NaN; 60 NaN particles after 4,000 steps
# of a thin slice of the column, tests/tools/recipe_stability.py).  SURVEY 8(d) says to raise the stiffness
# for blocks taller than ~1 m; the recipe now keeps the reference's own GAS_CONST
# (rho_pres_comp.glsl:33): sound speed 44.7 m/s, compression <= 4.4 %, no NaN in 3,250+ steps of the
# same slice, wall impact of the front included.
DAM_GAS_CONST = 2000.0


def dam_break_params(nx: int, ny: int, nz: int, spacing: float = 0.005) -> Params:
    s = np.float32(spacing)        # box faces in fp32 arithmetic: they define the grid, bit for bit
    p = Params()
    p.smoothing_coeff = 2.0
    p.mass = 1.2379e-4
    p.visc = 50.0
    p.gas_const = DAM_GAS_CONST
    p.gravity = [0.0, -9.80665, 0.0]
    p.dt = 1.0e-4
    lx, ly, lz = nx * s, ny * s, nz * s
    p.lower = [float(-s / 2), float(-s / 2), float(-s / 2), 1.0]
    p.upper = [float(3 * lx), float(2 * ly), float(lz + s / 2), 1.0]
    return p
