"""Hand-written onesweep radix sort (csrc/onesweep.cu) against numpy's stable sort: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(sph, keys, vals, key_bits):
    ko, vo = sph.binding.sort_pairs(keys, vals, key_bits)
    mask = np.uint32((1 << key_bits) - 1) if key_bits < 32 else np.uint32(0xFFFFFFFF)
    order = np.argsort(keys & mask, kind="stable")
    ref_vals = order.astype(np.uint32) if vals is None else vals[order]
    assert np.array_equal(ko, keys[order])
    assert np.array_equal(vo, ref_vals)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 4095, 4096, 4097, 8191, 100003, (1 << 20) + 5])
@pytest.mark.parametrize("key_bits", [1, 5, 8, 9, 16, 21, 24, 27, 32])
def test_sort_random(sph, n, key_bits):
    rng = np.random.default_rng(n * 131 + key_bits)
    hi = (1 << key_bits) - 1
    keys = rng.integers(0, hi, size=n, endpoint=True, dtype=np.uint64).astype(np.uint32)
    _check(sph, keys, None, key_bits)


def test_sort_with_values_and_high_garbage_bits(sph):
    rng = np.random.default_rng(7)
    n = 300001
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    vals = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    _check(sph, keys, vals, 13)      # bits above 13 must be ignored, order among equals stable
    _check(sph, keys, vals, 32)


@pytest.mark.parametrize("pattern", ["all_equal", "sorted", "reversed", "few_distinct", "all_ones"])
def test_sort_adversarial(sph, pattern):
    n = 1_000_003
    if pattern == "all_equal":
        keys = np.full(n, 12345, np.uint32)
    elif pattern == "sorted":
        keys = np.arange(n, dtype=np.uint32)
    elif pattern == "reversed":
        keys = np.arange(n, dtype=np.uint32)[::-1].copy()
    elif pattern == "few_distinct":
        keys = (np.arange(n, dtype=np.uint32) * 7919) % 3
    else:
        keys = np.full(n, 0xFFFFFFFF, np.uint32)
    _check(sph, keys, None, 32)
    _check(sph, keys, None, 20)


def test_sort_large(sph):
    rng = np.random.default_rng(99)
    n = 16_777_216
    keys = rng.integers(0, 12_582_913, size=n, dtype=np.uint64).astype(np.uint32)
    keys.sort()                                   # cell-ordered input, as in a running simulation
    keys[::97] = rng.integers(0, 12_582_913, size=len(keys[::97]), dtype=np.uint64).astype(np.uint32)
    _check(sph, keys, None, 24)


def test_sort_empty(sph):
    ko, vo = sph.binding.sort_pairs(np.zeros(0, np.uint32), None, 8)
    assert len(ko) == 0 and len(vo) == 0
