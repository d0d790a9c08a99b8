"""The CPU oracle against the golden vectors produced by executing the reference's own shader
text (tests/golden/make_golden.py).  Every pass is checked from IDENTICAL inputs (the golden
trajectory's own state), so the only differences are the oracle's canonical pow(q,3) = (q*q)*q
and pow(h,n) conventions versus the interpreter's correctly rounded pow (a few ulp)."""
import os

import numpy as np
import pytest

from conftest import assert_field_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ULP_TOL = 2e-6          # oracle canon vs correctly-rounded GLSL built-ins, per pass


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated")
    return np.load(path)


def _state(pos3, vel3, n):
    P = np.zeros((n, 16), np.float32)
    P[:, 0:3], P[:, 3], P[:, 4:7] = pos3, 1.0, vel3
    return P


def _check_step(oracle, g, s, P, p, sub):
    """P holds the golden input state of step s (pos, vel).  Runs the oracle pass by pass,
    feeding the GOLDEN output of each pass forward where the fixture has it for all particles."""
    oracle.pass_rho(P, p)
    rho_p = g[f"s{s}_rho_p"]
    assert_field_close(P[sub, 12], rho_p[:, 0], f"rho@{s}", tol=ULP_TOL)
    assert_field_close(P[sub, 13], rho_p[:, 1], f"pressure@{s}", tol=ULP_TOL,
                       scale=p.gas_const * (np.abs(rho_p[:, 0]) + p.resting_rho))
    if len(sub) == len(P):
        P[:, 12:14] = rho_p
    scale = oracle.force_scale(P, p)
    oracle.pass_force(P, p)
    assert_field_close(P[sub, 8:11], g[f"s{s}_force"], f"force@{s}", tol=ULP_TOL, scale=scale[sub])
    if len(sub) == len(P):
        P[:, 8:11] = g[f"s{s}_force"]
    vel_in = P[sub, 4:7].copy()
    oracle.pass_integrate(P, p)
    if len(sub) == len(P):      # identical inputs -> the integrate pass must be bit-exact
        assert np.array_equal(P[:, 0:3], g[f"s{s}_out_pos"], equal_nan=True)
        assert np.array_equal(P[:, 4:7], g[f"s{s}_out_vel"], equal_nan=True)
    else:
        dv_scale = np.abs(vel_in) + p.dt * scale[sub] / np.abs(P[sub, 12:13])
        assert_field_close(P[sub, 4:7], g[f"s{s}_out_vel"], f"vel@{s}", tol=ULP_TOL, scale=dv_scale)
        assert_field_close(P[sub, 0:3], g[f"s{s}_out_pos"], f"pos@{s}", tol=ULP_TOL)


def test_small_2000_all_walls(oracle):
    g = _load("small_2000.npz")
    n = int(g["num_particles"])
    p = oracle.default_params()
    p.mass, p.smoothing_coeff, p.visc, p.resting_rho = [float(x) for x in g["uniforms"]]
    for a in range(4):
        p.upper[a], p.lower[a] = float(g["upper"][a]), float(g["lower"][a])
    p.dt = np.float32(1.0) / np.float32(n)          # integrate_comp.glsl:33 with NUM_PARTICLES = n
    sub = np.arange(n)
    clamped = np.zeros(6, bool)
    for s in range(4):
        P = _state(g[f"s{s}_in_pos"], g[f"s{s}_in_vel"], n)
        _check_step(oracle, g, s, P, p, sub)
        out = g[f"s{s}_out_pos"]
        for a in range(3):
            clamped[a] |= (out[:, a] == np.float32(p.lower[a])).any()
            clamped[3 + a] |= (out[:, a] == np.float32(p.upper[a])).any()
    assert clamped.all(), f"fixture should hit all six walls, got {clamped}"


def test_default_10k_reference_scene(oracle):
    """BASELINE.json configs[0]: the repo-default 10,000-particle block, steps 0..2."""
    g = _load("default_10k.npz")
    p = oracle.default_params()
    sub = g["subset"]
    P0 = oracle.make_block(10, 100, 10)
    for s in range(3):
        if s == 0:
            P = P0.copy()
        else:
            P = _state(g[f"s{s}_in_pos"], g[f"s{s}_in_vel"], len(P0))
        _check_step(oracle, g, s, P, p, sub)
    w = g["w_lanes"]
    assert (w[0] == 1).all() and (w[1:] == 0).all(), "passes must not write .w lanes / age"


def test_nan_onset_matches_reference_text(oracle):
    g = _load("nan_onset.npz")
    p = oracle.default_params()
    P = oracle.make_block(10, 100, 10)
    first = -1
    for s in range(int(g["first_nan_step"]) + 2):
        oracle.step(P, p, 1)
        if np.isnan(P[:, :3]).any():
            first = s
            break
    assert first == int(g["first_nan_step"])
    assert np.array_equal(np.nonzero(np.isnan(P[:, :3]).any(axis=1))[0], g["nan_ids"])


def test_collider_extension_matches_its_frozen_fixture(oracle):
    """tests/golden/colliders.npz (make_colliders.py): 120 steps of a small block flying into a
    sphere and a box.  Not a pin against the reference (it has no obstacles) but against drift of
    the repo's own specification: the oracle must reproduce the frozen states bit for bit."""
    import importlib.util, os
    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_colliders", os.path.join(here, "make_colliders.py"))
    mc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mc)
    fx = np.load(os.path.join(here, "colliders.npz"))
    p, cols, P = mc.scene()
    assert np.array_equal(P.view(np.uint32), fx["start"].view(np.uint32))
    states = mc.run(p, cols, P.copy(), 120)
    assert np.array_equal(states.view(np.uint32), fx["states"].view(np.uint32))
    d = np.linalg.norm(states[-1][:, :3].astype(np.float64) - np.array(cols[0][1]), axis=1)
    assert abs(d.min() / cols[0][2] - 1.0) < 1e-6, "somebody should be sitting on the sphere"
