"""Parity of the CUDA path (through the C ABI) with the oracle.

Integer quantities (cell keys, slot permutation, cell table, neighbour counts) must be
bit-exact; fp32 fields must meet conftest.assert_field_close (<= 1e-5, the tolerance
BASELINE.json's north_star states); the integrate pass is bit-exact given the same inputs.
"""
import numpy as np
import pytest

from conftest import assert_field_close

pytestmark = pytest.mark.gpu

POS, VEL, FRC, RHO, PRS = slice(0, 3), slice(4, 7), slice(8, 11), 12, 13


def assert_bits_equal(got, want, msg):
    """Bit-for-bit equality; NaNs compare equal whatever their payload/sign bits."""
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan), msg + " (NaN pattern)"
    g = np.where(nan, np.float32(0), got).view(np.uint32)
    w = np.where(nan, np.float32(0), want).view(np.uint32)
    assert np.array_equal(g, w), msg


def make_sim(sph, p, cell_subdiv=1, counts=True):
    sim = sph.Simulation(cell_subdiv=cell_subdiv,
                         flags=sph.FLAG_COUNT_NEIGHBOURS if counts else 0)
    sim.apply_params(p)
    return sim


def check_grid_ints(sim, sph, oracle, P, p, cell_subdiv, prev_ids=None):
    """keys / permutation / cell table of the current arrangement vs the oracle's grid spec."""
    g = oracle.grid_setup(p, cell_subdiv)
    st = sim.stats()
    assert list(st.grid_dim) == list(g.dim) and st.num_cells == g.num_cells
    keys_by_id = oracle.cell_keys(P, g)
    sorted_keys = sim.debug_read(sph.DBG_SORTED_KEYS)
    ids = sim.debug_read(sph.DBG_SLOT_IDS)
    if prev_ids is None:
        prev_ids = np.arange(len(P), dtype=np.uint32)
    perm_want = np.argsort(keys_by_id[prev_ids], kind="stable").astype(np.uint32)
    assert np.array_equal(sim.debug_read(sph.DBG_LAST_PERM), perm_want)
    assert np.array_equal(ids, prev_ids[perm_want])
    assert np.array_equal(sorted_keys, keys_by_id[ids])
    cs = sim.debug_read(sph.DBG_CELL_START)
    want = np.searchsorted(sorted_keys, np.arange(g.num_cells + 2, dtype=np.uint64), side="left")
    assert np.array_equal(cs, want.astype(np.uint32))
    return ids


def check_passes(sim, sph, oracle, P, p, cell_subdiv, gpu_oracle=False):
    """One step pass by pass from the host state P (oracle advances P in place)."""
    sim.upload(P)
    sim.pass_rho()
    ids = check_grid_ints(sim, sph, oracle, P, p, cell_subdiv)
    if gpu_oracle:
        c_rho = oracle.gpu_pass(0, P, p, counts=True)
    else:
        c_rho = oracle.pass_rho(P, p, counts=True)
    G = sim.download()
    assert np.array_equal(sim.debug_read(sph.DBG_COUNTS_RHO), c_rho)
    assert_field_close(G[:, RHO], P[:, RHO], "rho")
    assert_field_close(G[:, PRS], P[:, PRS], "pressure",
                       scale=p.gas_const * (np.abs(P[:, RHO]) + p.resting_rho))
    assert_bits_equal(G[:, POS], P[:, POS], "rho pass must not touch pos")
    assert_bits_equal(G[:, VEL], P[:, VEL], "rho pass must not touch vel")

    # force pass on IDENTICAL inputs: hand the oracle's rho/p to the GPU
    sim.upload(P)
    sim.pass_force()
    if gpu_oracle:
        scale = oracle.gpu_force_scale(P, p)
        c_f = oracle.gpu_pass(1, P, p, counts=True)
    else:
        scale = oracle.force_scale(P, p)
        c_f = oracle.pass_force(P, p, counts=True)
    G = sim.download()
    assert np.array_equal(sim.debug_read(sph.DBG_COUNTS_FORCE), c_f)
    assert_field_close(G[:, FRC], P[:, FRC], "force", scale=scale)

    # integrate on identical inputs: bit-exact
    sim.upload(P)
    sim.pass_integrate()
    oracle.pass_integrate(P, p)
    G = sim.download()
    assert_bits_equal(G, P, "integrate must be bit-exact")
    return ids


@pytest.mark.parametrize("cell_subdiv", [1, 2, 4])
def test_config1_default_scene_pass_by_pass(sph, oracle, cell_subdiv):
    """BASELINE.json configs[0]: the reference's 10x100x10 block with its shader constants."""
    p = oracle.default_params()
    sim = make_sim(sph, p, cell_subdiv)
    P = oracle.make_block(10, 100, 10)
    assert np.array_equal(sim.download(), P), "default scene must equal make_grid()/init_particles()"
    for _ in range(3):
        check_passes(sim, sph, oracle, P, p, cell_subdiv)


def test_config1_steps_resynced_through_nan_onset(sph, oracle):
    """Per-step parity on identical inputs for 60 steps, including the NaN onset at step 50
    (two particles clamped onto the same corner -> normalize(0))."""
    p = oracle.default_params()
    sim = make_sim(sph, p, 1, counts=False)
    sim.set_paused(False)
    P = oracle.make_block(10, 100, 10)
    first_nan = None
    chk = make_sim(sph, p, 1, counts=True)
    for s in range(60):
        # (a) every pass on identical inputs, element-wise gates included
        check_passes(chk, sph, oracle, P.copy(), p, 1)
        # (b) the whole step through nprsph_step: norm gates (later passes see the earlier
        #     passes' rounding, so element-wise conditioning no longer applies)
        sim.upload(P)
        sim.step(1)
        oracle.step(P, p, 1)
        G = sim.download()
        assert_field_close(G[:, RHO], P[:, RHO], f"rho@{s}")
        assert_field_close(G[:, FRC], P[:, FRC], f"force@{s}", elementwise=False)
        assert_field_close(G[:, VEL], P[:, VEL], f"vel@{s}", elementwise=False)
        assert_field_close(G[:, POS], P[:, POS], f"pos@{s}", elementwise=False)
        if first_nan is None and np.isnan(P[:, POS]).any():
            first_nan = s
    assert first_nan == 50
    assert sim.stats().nan_particles == np.isnan(P[:, POS]).any(axis=1).sum()


@pytest.mark.parametrize("cell_subdiv", [1, 2, 3])
def test_dam_break_small_free_running(sph, oracle, cell_subdiv):
    """Stable dam-break recipe, 20x24x16 block, free running for 200 steps: the trajectory
    drift against the oracle stays bounded (reported in units of h)."""
    nx, ny, nz = 20, 24, 16
    p = oracle.dam_break_params(nx, ny, nz)
    P = oracle.jitter(oracle.make_block(nx, ny, nz), 1e-4 * 0.005, seed=1234)
    sim = make_sim(sph, p, cell_subdiv, counts=False)
    sim.upload(P)
    sim.set_paused(False)
    h = float(oracle.smoothing_length(p))
    for s in range(8):
        sim.step(25)
        oracle.step(P, p, 25, grid=1)
        G = sim.download()
        drift = np.abs(G[:, POS].astype(np.float64) - P[:, POS]).max() / h
        assert drift < 1e-2, f"after {25 * (s + 1)} steps: drift {drift:.3e} h"
        keys = sim.debug_read(sph.DBG_SORTED_KEYS)
        ids = sim.debug_read(sph.DBG_SLOT_IDS)
        assert (np.diff(keys.astype(np.int64)) >= 0).all()
        assert np.array_equal(np.sort(ids), np.arange(len(P), dtype=np.uint32))
    assert not np.isnan(G).any()
    print(f"\n[drift] cell_subdiv={cell_subdiv}: max |x_gpu - x_oracle| / h after 200 steps = {drift:.3e}")


def test_long_axis_far_from_the_origin(sph, oracle):
    """A 16000-cell axis with the fluid 80 m from the grid origin: cell keys / permutation / cell
    table and neighbour counts still match the oracle bit for bit (the cell coordinate is fp64 on
    both sides; the cell size no longer grows with the grid), fields to the usual tolerance."""
    p = oracle.dam_break_params(12, 14, 10)
    p.lower[0], p.upper[0] = -40.0, 40.0
    P = oracle.jitter(oracle.make_block(12, 14, 10), 0.4 * 0.005, seed=13)
    P[:, 0] += np.float32(39.9)
    sim = make_sim(sph, p, cell_subdiv=2)
    for _ in range(3):
        check_passes(sim, sph, oracle, P, p, 2)
    assert sim.stats().grid_dim[0] > 14000


def test_iterated_permutation_every_step(sph, oracle):
    """Slot permutation == iterated stable sort of the oracle's keys, checked on every step."""
    nx, ny, nz = 12, 20, 10
    p = oracle.dam_break_params(nx, ny, nz)
    p.gravity[1] = -200.0                      # make particles cross cells quickly
    P = oracle.jitter(oracle.make_block(nx, ny, nz), 0.3 * 0.005, seed=5)
    sim = make_sim(sph, p, 1)
    sim.upload(P)
    sim.set_paused(False)
    ids = None
    moved = 0
    for s in range(40):
        G = sim.download()
        new_ids = check_grid_ints(sim, sph, oracle, G, p, 1, ids)
        if ids is not None:
            moved += int((new_ids != ids).sum())
        ids = new_ids
        sim.step(1)
    assert moved > 0, "scene too static to exercise the re-sort"


@pytest.mark.parametrize("cell_subdiv", [1, 2, 3, 4])
def test_ragged_sizes_and_degenerate_inputs(sph, oracle, cell_subdiv):
    p = oracle.dam_break_params(10, 10, 10)
    rng = np.random.default_rng(3)
    for n in (1, 2, 33, 257, 4097):
        P = np.zeros((n, 16), np.float32)
        P[:, POS] = rng.uniform(0.0, 0.05, size=(n, 3)).astype(np.float32)
        P[:, 3] = 1.0
        P[:, VEL] = rng.normal(0, 0.1, size=(n, 3)).astype(np.float32)
        sim = make_sim(sph, p, cell_subdiv)
        check_passes(sim, sph, oracle, P, p, cell_subdiv)
    # coincident particles, NaN / Inf positions, particles outside the box
    n = 600
    P = np.zeros((n, 16), np.float32)
    P[:, POS] = rng.uniform(0.0, 0.04, size=(n, 3)).astype(np.float32)
    P[:, 3] = 1.0
    P[10, POS] = P[11, POS]                      # r == 0 between distinct particles -> NaN force
    P[20, 0] = np.nan
    P[21, 1] = np.inf
    P[22, POS] = (-5.0, 7.0, 0.01)               # far outside the box
    P[23, POS] = (-5.0, 7.0, 0.0105)             # ... with a neighbour out there
    sim = make_sim(sph, p, cell_subdiv)
    check_passes(sim, sph, oracle, P, p, cell_subdiv)
    assert np.isnan(P[10, POS]).all() and np.isnan(P[20, POS]).any()


def test_w_lanes_and_age_survive(sph, oracle):
    """The shaders never write .w lanes or extras[2..3]."""
    p = oracle.dam_break_params(8, 8, 8)
    P = oracle.make_block(8, 8, 8)
    rng = np.random.default_rng(11)
    for col in (3, 7, 11, 14, 15):
        P[:, col] = rng.normal(size=len(P)).astype(np.float32)
    sim = make_sim(sph, p, 1, counts=False)
    sim.upload(P)
    sim.set_paused(False)
    sim.step(5)
    G = sim.download()
    for col in (3, 7, 11, 14, 15):
        assert np.array_equal(G[:, col], P[:, col]), col
    assert not np.array_equal(G[:, POS], P[:, POS])


def test_empty_buffer(sph):
    sim = sph.Simulation()
    sim.upload(np.zeros((0, 16), np.float32))
    sim.set_paused(False)
    sim.step(3)
    assert sim.num_particles == 0 and sim.download().shape == (0, 16)
    assert sim.stats().nan_particles == 0


def test_config2_one_million_vs_all_pairs(sph, oracle):
    """BASELINE.json configs[1]: 1M-particle dam break, uniform grid vs brute-force all-pairs
    (the CUDA all-pairs oracle is first checked bit-for-bit against the C oracle at 20k)."""
    small_p = oracle.dam_break_params(20, 40, 25)
    S = oracle.jitter(oracle.make_block(20, 40, 25), 0.2 * 0.005, seed=9)
    S2 = S.copy()
    c1 = oracle.pass_rho(S, small_p, counts=True)
    c2 = oracle.gpu_pass(0, S2, small_p, counts=True)
    assert np.array_equal(c1, c2) and np.array_equal(S.view(np.uint32), S2.view(np.uint32))
    c1 = oracle.pass_force(S, small_p, counts=True)
    c2 = oracle.gpu_pass(1, S2, small_p, counts=True)
    assert np.array_equal(c1, c2) and np.array_equal(S.view(np.uint32), S2.view(np.uint32))

    n1 = 100
    p = oracle.dam_break_params(n1, n1, n1)
    P = oracle.jitter(oracle.make_block(n1, n1, n1), 0.05 * 0.005, seed=1234)
    sim = make_sim(sph, p, 2)
    # advance a few steps on the GPU first so the state is not a lattice
    sim.upload(P)
    sim.set_paused(False)
    sim.step(5)
    P = sim.download()
    check_passes(sim, sph, oracle, P, p, 2, gpu_oracle=True)
    # the grid-accelerated C oracle must agree with brute force bit for bit as well
    Q = P.copy(); R = P.copy()
    cq = oracle.pass_rho(Q, p, counts=True, grid=1)
    cr = oracle.gpu_pass(0, R, p, counts=True)
    assert np.array_equal(cq, cr) and np.array_equal(Q.view(np.uint32), R.view(np.uint32))


def test_config3_sixteen_million_full_size(sph, oracle):
    """BASELINE.json configs[2] at its full size (256^3 = 16,777,216 particles, the bench
    workload): every pass of one step against the oracle's grid-accelerated variant, which is
    bit-identical to the all-pairs loop (checked above at 1M).  Integers exact (keys, sortedness,
    permutation, neighbour counts incl./excl. self), fp32 fields at 1e-5 (norm gates for the
    force: its conditioning scale is an all-pairs sum), integrate bit-exact; plus the
    size-independent properties: counts are symmetric (an even total) and the slot ids are a
    permutation."""
    n1 = 256
    p = oracle.dam_break_params(n1, n1, n1)
    sim = make_sim(sph, p, 2)
    sim.scene_block(n1, n1, n1, 0.005, None, 1e-4 * 0.005, 1234)
    sim.set_paused(False)
    sim.step(3)                                   # off the lattice, arrangement no longer the identity
    P = sim.download()
    n = len(P)
    assert n == 16_777_216 and not np.isnan(P).any()

    sim.upload(P)
    sim.pass_rho()
    g = oracle.grid_setup(p, 2)
    keys = oracle.cell_keys(P, g)
    ids = sim.debug_read(sph.DBG_SLOT_IDS)
    sk = sim.debug_read(sph.DBG_SORTED_KEYS)
    assert np.array_equal(np.bincount(ids, minlength=n), np.ones(n, np.int64)), "slot ids: a permutation"
    assert np.array_equal(sk, keys[ids]) and np.all(sk[1:] >= sk[:-1])
    same = sk[1:] == sk[:-1]
    assert np.all(ids[1:][same] > ids[:-1][same]), "stable: equal keys keep their upload order"

    Q = P.copy()
    c_rho = oracle.pass_rho(Q, p, counts=True, grid=2)
    G = sim.download()
    assert np.array_equal(sim.debug_read(sph.DBG_COUNTS_RHO), c_rho)
    assert_field_close(G[:, RHO], Q[:, RHO], "rho@16M")
    assert_field_close(G[:, PRS], Q[:, PRS], "pressure@16M",
                       scale=p.gas_const * (np.abs(Q[:, RHO]) + p.resting_rho))

    sim.upload(Q)                                 # identical inputs for the force pass
    sim.pass_force()
    c_f = oracle.pass_force(Q, p, counts=True, grid=2)
    G = sim.download()
    got = sim.debug_read(sph.DBG_COUNTS_FORCE)
    assert np.array_equal(got, c_f)
    assert np.array_equal(c_f + 1, c_rho), "density counts the particle itself, the force pass does not"
    assert int(c_f.sum(dtype=np.int64)) % 2 == 0, "the neighbour relation is symmetric"
    assert_field_close(G[:, FRC], Q[:, FRC], "force@16M", elementwise=False)

    sim.upload(Q)
    sim.pass_integrate()
    oracle.pass_integrate(Q, p)
    assert_bits_equal(sim.download(), Q, "integrate must be bit-exact @16M")


def step_vs_oracle(sim, sph, oracle, P, p, grid, tag):
    """ONE nprsph_step (column records, deferred queues, fused force + integrate) from the host state
    P, checked pass by pass on the step's OWN intermediate results: the oracle's density pass on P;
    its force pass on P carrying the GPU's density (a stiff equation of state turns a 1e-5 density
    difference into a 5e-4 pressure difference, so a force computed from the oracle's density is not
    comparable); its integrate pass on the GPU's force, which must then match bit for bit."""
    sim.upload(P)
    sim.set_paused(False)
    sim.step(1)
    G = sim.download()
    Q = P.copy()
    c_rho = oracle.pass_rho(Q, p, counts=True, grid=grid)
    assert np.array_equal(sim.debug_read(sph.DBG_COUNTS_RHO), c_rho), f"{tag}: density neighbour counts"
    assert_field_close(G[:, RHO], Q[:, RHO], f"rho {tag}")
    assert_field_close(G[:, PRS], Q[:, PRS], f"pressure {tag}", scale=p.gas_const * (np.abs(Q[:, RHO]) + p.resting_rho))
    Q[:, RHO], Q[:, PRS] = G[:, RHO], G[:, PRS]            # identical inputs for the force pass
    scale = oracle.force_scale(Q, p, grid=grid)
    c_f = oracle.pass_force(Q, p, counts=True, grid=grid)
    assert np.array_equal(sim.debug_read(sph.DBG_COUNTS_FORCE), c_f), f"{tag}: force neighbour counts"
    assert_field_close(G[:, FRC], Q[:, FRC], f"force {tag}", scale=scale)
    Q[:, FRC] = G[:, FRC]                                  # identical inputs for the integrate pass
    oracle.pass_integrate(Q, p)
    assert_bits_equal(G, Q, f"{tag}: the fused integrate epilogue must be bit-exact")
    return G


@pytest.mark.parametrize("cell_subdiv", [0, 3, 4])
def test_default_scene_whole_steps_replay_records_at_the_reference_smoothing(sph, oracle, cell_subdiv):
    """BASELINE configs[0] through nprsph_step: at the reference's own h = 4 lattice spacings the
    automatic grid (cell = h/4, 81 columns per walk) keeps every walk inside the column-record
    format, so the force pass replays records instead of re-testing candidates."""
    p = oracle.default_params()
    sim = make_sim(sph, p, cell_subdiv)
    assert sim.stats().cell_subdiv == (cell_subdiv or 4)
    P = oracle.make_block(10, 100, 10)
    for s in range(4):
        P = step_vs_oracle(sim, sph, oracle, P, p, 1, f"default scene step {s} subdiv {cell_subdiv}")
    if cell_subdiv in (0, 4):
        st = sim.walk_stats()
        assert st["pair_walks"] * 2 > 0.95 * len(P), st        # the lattice pairs up; almost nothing is deferred


def test_config3_sixteen_million_evolved_whole_step(sph, oracle):
    """BASELINE configs[2] AFTER the dam has broken (2,000 steps): one whole nprsph_step -- column
    records, the deferred-slot kernels that carry a few percent of a disordered fluid, fused force +
    integrate -- against the oracle's grid variant; the recipe must also have kept the 1.28 m column
    free of NaNs."""
    n1 = 256
    p = oracle.dam_break_params(n1, n1, n1)
    sim = make_sim(sph, p, 2)
    sim.scene_block(n1, n1, n1, 0.005, None, 1e-4 * 0.005, 1234)
    sim.set_paused(False)
    sim.step(2000)
    P = sim.download()
    # The reference's boundary rule parks particles exactly ON the walls, and wall particles carry no
    # pressure (rho < rho0 -> p = 0), so the bottom corners of the block collect particles: two on the
    # same corner are coincident -> normalize(0) = NaN (the default scene does
    # it at step 50).  Stiffness does not change that (scripts/nan_probe.py: 4 NaN particles at step
    # 1250 for any wall gap / gas_const); what the recipe must prevent is the bulk blow-up.
    nan = np.isnan(P[:, POS]).any(axis=1)
    assert nan.sum() <= 16, f"{nan.sum()} NaN particles: the dam-break recipe is unstable"
    ok = ~nan
    assert P[ok, RHO].max() < 1.15 * p.resting_rho, "compression of the 1.28 m column must stay below 15 %"
    st = sim.walk_stats()
    assert st["single_walks"] > 0, "an evolved fluid has slots that cannot share a walk (deferred queue)"
    step_vs_oracle(sim, sph, oracle, P, p, 2, "@16M evolved")
    print(f"\n[evolved 16M] {st['single_walks']} deferred slots, {st['distance_tests'] / len(P):.1f} tests and "
          f"{st['neighbours'] / len(P):.1f} neighbours per particle")
