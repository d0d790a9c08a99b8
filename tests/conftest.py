import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ---- the parity tolerance (BASELINE.json north_star: fp32 fields within 1e-5 per step) ----------
REL_TOL = 1e-5


def assert_field_close(got, want, name, tol=REL_TOL, scale=None, elementwise=True):
    """NaN-aware comparison of one fp32 field of all particles (shape (n,) or (n, k)).

    SPH sums cancel (opposing neighbours), so element-wise relative error is meaningless near
    equilibrium; the gate is:
      * identical NaN / Inf pattern,
      * max|got - want| <= tol * max|want|                      (norm-relative),
      * |got - want| <= tol * (|want| + rms(want)) element-wise, or, when `scale` is given
        (oracle.force_scale: the sum of |terms| behind each force component, i.e. the
        conditioning of that sum), |got - want| <= tol * scale element-wise.
    `elementwise=False` keeps the two norm gates only (whole-step checks, where the inputs of
    the later passes already differ by the earlier passes' rounding).
    """
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, name
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{name}: NaN pattern differs ({nan_g.sum()} vs {nan_w.sum()})"
    inf_w = np.isinf(want)
    assert np.array_equal(got[inf_w], want[inf_w]), f"{name}: Inf pattern differs"
    ok = ~(nan_w | inf_w)
    if not ok.any():
        return 0.0
    g, w = got[ok], want[ok]
    scale = np.abs(w).max()
    err = np.abs(g - w)
    if scale == 0.0:
        assert err.max() == 0.0, f"{name}: expected all zeros, max err {err.max()}"
        return 0.0
    rel = err.max() / scale
    assert rel <= tol, f"{name}: norm-relative error {rel:.3e} > {tol}"
    l2 = np.sqrt(np.sum(err * err) / np.sum(w * w))
    assert l2 <= tol, f"{name}: L2-relative error {l2:.3e} > {tol}"
    if not elementwise:
        return rel
    if scale is not None:
        bound = tol * np.maximum(np.asarray(scale, np.float64)[ok], np.abs(w))
    else:
        rms = np.sqrt(np.mean(w * w))
        bound = tol * (np.abs(w) + rms)
    bound = np.maximum(bound, np.finfo(np.float32).tiny)
    worst = (err / bound).max()
    assert worst <= 1.0, f"{name}: element-wise error {worst:.3f}x the bound"
    return rel


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def sph():
    import nprsph_b200
    nprsph_b200.load()
    return nprsph_b200
