"""Self-consistency of the CPU oracle (no GPU): known answers of SURVEY.md Appendix C, the
grid-accelerated variant against the literal all-pairs loops, and the sqrt-free predicate."""
import numpy as np
import pytest


def test_appendix_c_known_answers(oracle):
    p = oracle.default_params()
    P = oracle.make_block(10, 100, 10)
    c = oracle.pass_rho(P, p, counts=True)
    rho = P[:, 12]
    assert (c.min(), c.max()) == (51, 253) and c.mean() == pytest.approx(181.2, abs=0.05)
    assert rho.min() == pytest.approx(44329.7, rel=1e-6)
    assert rho.mean() == pytest.approx(130256.5, rel=1e-6)
    assert rho.max() == pytest.approx(160000.0, rel=1e-6)
    oracle.pass_force(P, p)
    assert P[:, 13].max() == pytest.approx(3.18e8, rel=1e-3)
    assert np.abs(P[:, 8:11]).max() == pytest.approx(1.99e10, rel=1e-2)
    oracle.pass_integrate(P, p)
    assert np.abs(P[:, 4:7]).max() == pytest.approx(20.6, rel=1e-2)


def test_first_nan_at_step_50(oracle):
    p = oracle.default_params()
    P = oracle.make_block(10, 100, 10)
    for s in range(51):
        assert not np.isnan(P).any(), s
        oracle.step(P, p, 1)
    bad = np.nonzero(np.isnan(P[:, :3]).any(axis=1))[0]
    assert list(bad) == [0, 10]


@pytest.mark.parametrize("cell_subdiv", [1, 2, 3])
def test_grid_variant_is_bit_identical_to_all_pairs(oracle, cell_subdiv):
    p = oracle.dam_break_params(12, 16, 10)
    A = oracle.jitter(oracle.make_block(12, 16, 10), 0.4 * 0.005, seed=77)
    A[5, 0] = np.nan
    A[6, :3] = A[7, :3]
    B = A.copy()
    for _ in range(3):
        ca = oracle.pass_rho(A, p, counts=True)
        cb = oracle.pass_rho(B, p, counts=True, grid=cell_subdiv)
        assert np.array_equal(ca, cb)
        ca = oracle.pass_force(A, p, counts=True)
        cb = oracle.pass_force(B, p, counts=True, grid=cell_subdiv)
        assert np.array_equal(ca, cb)
        oracle.pass_integrate(A, p)
        oracle.pass_integrate(B, p)
        assert np.array_equal(A, B, equal_nan=True)


@pytest.mark.parametrize("cell_subdiv", [1, 2])
def test_reach_suffices_on_a_16000_cell_axis(oracle, cell_subdiv):
    """The cell is only 2^-14 wider than h / subdiv and the cell coordinate is computed in fp64, so
    `reach` cells must still reach every neighbour when an axis holds ~16000 cells and the fluid
    sits 80 m from the grid origin (fp32 positions there have a 4e-6 ulp, 1/1250 of the spacing):
    the grid variant stays bit-identical to the all-pairs loops."""
    p = oracle.dam_break_params(10, 12, 8)
    p.lower[0], p.upper[0] = -40.0, 40.0
    g = oracle.grid_setup(p, cell_subdiv)
    assert g.dim[0] > 7000 * cell_subdiv
    A = oracle.jitter(oracle.make_block(10, 12, 8), 0.4 * 0.005, seed=9)
    A[:, 0] += np.float32(39.9)
    B = A.copy()
    for _ in range(3):
        ca = oracle.pass_rho(A, p, counts=True)
        cb = oracle.pass_rho(B, p, counts=True, grid=cell_subdiv)
        assert np.array_equal(ca, cb) and ca.min() > 5
        ca = oracle.pass_force(A, p, counts=True)
        cb = oracle.pass_force(B, p, counts=True, grid=cell_subdiv)
        assert np.array_equal(ca, cb)
        oracle.pass_integrate(A, p)
        oracle.pass_integrate(B, p)
        assert np.array_equal(A, B, equal_nan=True)


def test_r2_threshold_is_the_sqrt_predicate(oracle):
    """(sqrtf(r2) < h) == (r2 < T) for every fp32 r2 around the threshold."""
    for h in (0.02, 0.01, 0.0125, 0.035, 1.0, 3e-5, 7.5):
        h = np.float32(h)
        T = oracle.r2_threshold(h)
        r2 = np.float32(T)
        lo = r2
        for _ in range(200):
            lo = np.nextafter(lo, np.float32(0))
        xs = [lo]
        for _ in range(400):
            xs.append(np.nextafter(xs[-1], np.float32(np.inf)))
        xs = np.array(xs, np.float32)
        assert np.array_equal(np.sqrt(xs) < h, xs < T)


def test_cell_keys_cover_every_neighbour(oracle):
    """A pair that passes the predicate is never more than `reach` cells apart on any axis."""
    p = oracle.dam_break_params(10, 10, 10)
    rng = np.random.default_rng(1)
    P = np.zeros((4000, 16), np.float32)
    P[:, :3] = rng.uniform(-0.0025, 0.05, (4000, 3)).astype(np.float32)
    h = float(oracle.smoothing_length(p))
    for k in (1, 2):
        g = oracle.grid_setup(p, k)
        keys = oracle.cell_keys(P, g).astype(np.int64)
        cz = keys % g.dim[2]; cy = (keys // g.dim[2]) % g.dim[1]; cx = keys // (g.dim[2] * g.dim[1])
        d = P[:, None, :3] - P[None, :1000, :3]
        r = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])
        i, j = np.nonzero(r < np.float32(h))
        for c in (cx, cy, cz):
            assert np.abs(c[i] - c[j]).max() <= k


def test_collider_spec_of_the_oracle(oracle):
    """Static colliders are an extension (README.md:59 future work); the oracle is their
    specification.  A particle that ends a step inside an obstacle is put on its surface and its
    normal velocity is multiplied by -damping, like the box walls (integrate_comp.glsl:46-77)."""
    p = oracle.dam_break_params(4, 4, 4)
    p.dt = 1e-3
    P = np.zeros((4, 16), np.float32)
    P[:, 3] = 1.0
    P[:, 12] = 1000.0                                       # rho (force stays 0: pure advection)
    P[0, :3] = (0.010, 0.010, 0.0055); P[0, 4:7] = (0.0, 0.0, 1.0)     # flies into the sphere from -z
    P[1, :3] = (0.010, 0.010, 0.010); P[1, 4:7] = (0.0, 0.0, 0.0)     # sits at the sphere's centre
    P[2, :3] = (0.0301, 0.010, 0.010); P[2, 4:7] = (-1.0, 0.5, 0.0)   # enters the box through its +x face
    P[3, :3] = (0.015, 0.018, 0.010); P[3, 4:7] = (0.1, 0.1, 0.1)     # touches nothing
    cs = [oracle.sphere((0.010, 0.010, 0.010), 0.004), oracle.box((0.020, 0.0, 0.0), (0.030, 0.02, 0.02))]
    Q = P.copy()
    oracle.pass_integrate(P, p, cs)
    oracle.pass_integrate(Q, p)                              # no colliders
    c, R = np.float32([0.010, 0.010, 0.010]), np.float32(0.004)
    # 0: on the sphere's surface below the centre, vz reflected and damped
    assert np.linalg.norm(P[0, :3] - c) == pytest.approx(R, rel=1e-6) and P[0, 2] < c[2]
    assert P[0, 6] == pytest.approx(-p.damping * 1.0, rel=1e-6) and P[0, 4] == 0 and P[0, 5] == 0
    # 1: pushed out along +y
    assert tuple(P[1, :3]) == (c[0], np.float32(c[1] + R), c[2])
    # 2: back on the +x face, vx reflected, vy untouched
    assert P[2, 0] == np.float32(0.030) and P[2, 4] == np.float32(p.damping) and P[2, 5] == np.float32(0.5)
    # 3: identical to the collider-free pass
    assert np.array_equal(P[3], Q[3])


def test_package_scene_recipes_equal_the_oracles(oracle):
    """bench.py configures the product from npr-sph_b200/scenes.py (no oracle import on the product
    arm); the parity tests use the oracle's copy.  The two must describe the same scenes."""
    import nprsph_b200 as sph
    pairs = [(sph.scenes.default_params(), oracle.default_params())]
    for dims in ((256, 256, 256), (20, 24, 16), (512, 256, 512)):
        pairs.append((sph.scenes.dam_break_params(*dims), oracle.dam_break_params(*dims)))
    for a, b in pairs:
        for k in ("mass", "smoothing_coeff", "visc", "resting_rho", "particle_radius", "gas_const", "damping", "dt", "pi"):
            assert np.float32(getattr(a, k)) == np.float32(getattr(b, k)), k
        for i in range(3):
            assert np.float32(a.gravity[i]) == np.float32(b.gravity[i])
            assert np.float32(a.upper[i]) == np.float32(b.upper[i]) and np.float32(a.lower[i]) == np.float32(b.lower[i])
