"""Two-sided comparators (the reference's own agreesWith is one-sided and NaN-blind,
tests/test_utilities.hpp:470-487)."""
import numpy as np

TOL = 1e-12          # BASELINE.json north_star: max |delta amp| <= 1e-12, relative expectation error <= 1e-12


def max_abs_diff(a, b):
    a = np.asarray(a, dtype=np.complex128).reshape(-1)
    b = np.asarray(b, dtype=np.complex128).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.abs(a - b)
    assert not np.isnan(d).any(), "NaN in comparison"
    return float(d.max()) if d.size else 0.0


def assert_close(got, want, tol=TOL, what=""):
    """max |delta| <= tol * max(1, max|want|): amplitudes are O(1) in the reference's tests, random
    non-unitary gates can scale them up, so the bound is relative to the largest reference amplitude."""
    scale = max(1.0, float(np.abs(np.asarray(want)).max()))
    d = max_abs_diff(got, want)
    assert d <= tol * scale, "%s max|delta|=%.3e > %.1e*%.3g" % (what, d, tol, scale)


def assert_exact(got, want, what=""):
    """Value equality on every double (+0 == -0): permutation / sign-only ops must be bit-exact."""
    got = np.asarray(got, dtype=np.complex128).reshape(-1)
    want = np.asarray(want, dtype=np.complex128).reshape(-1)
    assert got.shape == want.shape
    bad = np.nonzero((got.real != want.real) | (got.imag != want.imag))[0]
    assert bad.size == 0, "%s differs at %d of %d amplitudes (first %d: %r vs %r)" % (what, bad.size, got.size, bad[0], got[bad[0]], want[bad[0]])


def assert_value_close(got, want, tol=TOL, what=""):
    got, want = complex(got), complex(want)
    assert abs(got - want) <= tol * max(1.0, abs(want)), "%s %r vs %r" % (what, got, want)
