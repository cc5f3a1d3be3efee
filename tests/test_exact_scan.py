"""oracle/exact_scan.py: the parallel-friendly scan reproduces the reference's serial Float64 cumsum (resample.jl:18-21) bit for
bit, with a sequential part of a few dozen elements — on normalised particle weights, adversarial ties, binade crossings and
degenerate weights (groundwork for DESIGN.md §11 item 1; not used by the product)."""
import numpy as np
import pytest

from oracle import exact_scan as X


def _weights(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        w = np.full(n, 1.0 / n)
    elif kind == "softmax":
        lw = rng.standard_normal(n) * 2.0
        w = np.exp(lw - lw.max()); w /= w.sum()
    elif kind == "degenerate":
        lw = rng.standard_normal(n) * 30.0
        w = np.exp(lw - lw.max()); w /= w.sum()
    elif kind == "ties":                      # weights whose low bits make exact half-ulp ties against the running sum
        w = np.full(n, 2.0 ** -12)
        w[::7] += 2.0 ** -54                  # = half an ulp of sums in [0.5, 1)
        w[3::11] += 2.0 ** -55
    elif kind == "dyadic":
        w = 2.0 ** -rng.integers(8, 30, n).astype(np.float64)
    else:
        raise ValueError(kind)
    return w


@pytest.mark.parametrize("kind", ["uniform", "softmax", "degenerate", "ties", "dyadic"])
@pytest.mark.parametrize("n", [1, 2, 37, 1000, 20000])
def test_exact_scan_equals_serial_cumsum_bit_for_bit(kind, n):
    w = _weights(kind, n, seed=n)
    ref = X.serial_cumsum(w)
    got, st = X.exact_scan(w, return_stats=True)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (kind, n, int(np.argmax(got != ref)))
    if kind in ("uniform", "softmax") and n >= 1000:
        assert st["special"] <= 120, st       # the sequential part: binade crossings + the start + a few ties


def test_fixed_point_scan_is_not_bit_identical_but_the_exact_scan_is():
    """what the product's FAST scan returns (exact sums of the weights quantised to 2^-62) differs from the serial cumsum in the
    last bits for generic weights — the reason for the FAST-mode caveat of DESIGN.md §6"""
    w = _weights("softmax", 50000, seed=3)
    ref = X.serial_cumsum(w)
    q = np.round(w * 2.0 ** 62).astype(object)
    fast = np.array([float(v) for v in np.cumsum(q)], dtype=np.float64) * 2.0 ** -62
    assert (fast != ref).mean() > 0.05
    assert np.abs(fast - ref).max() < 1e-11
    assert np.array_equal(X.exact_scan(w), ref)


@pytest.mark.parametrize("kind", ["uniform", "softmax", "degenerate", "ties", "dyadic"])
@pytest.mark.parametrize("n,chunk", [(1, 8), (37, 8), (1000, 64), (20000, 3544), (20000, 257)])
def test_blocked_decomposition_equals_serial_cumsum_bit_for_bit(kind, n, chunk):
    """the same algorithm in the decomposition a GPU would run: block-local scans, scans over block aggregates, one walk over
    the specials — independent of the block size"""
    w = _weights(kind, n, seed=n + 1)
    ref = X.serial_cumsum(w)
    got, st = X.exact_scan_blocked(w, chunk=chunk, return_stats=True)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (kind, n, chunk, int(np.argmax(got != ref)))
    assert np.array_equal(got, X.exact_scan(w))
