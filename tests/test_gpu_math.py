"""Accuracy of the kernels' FP64 primitives (fast_rcp, sincos_flow) against NumPy/glibc, in ulps."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ulps(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_fast_rcp_within_one_ulp(handle):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(1e-3, 1e5, 400000), -rng.uniform(1e-3, 1e5, 100000),
                        10.0 ** rng.uniform(-30, 30, 100000), [1.0, 2.0, 3.0, 0.1, 1e-300, 1e300]])
    r, _, _ = handle.math_probe(x)
    u = _ulps(r, 1.0 / x)
    ok = np.abs(x) > 1e-290         # the approximation flushes subnormal results (never reached by r (r-rs) sin th)
    assert u[ok].max() <= 1.0, u[ok].max()
    assert (u[ok] == 0).mean() > 0.85          # correctly rounded most of the time


def test_sincos_flow_within_two_ulps(handle):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(0.0, np.pi, 400000), rng.uniform(-50.0, 50.0, 200000),
                        rng.uniform(-1e4, 1e4, 50000),
                        [0.0, np.pi / 2, np.pi, np.pi / 4, 3 * np.pi / 4, -np.pi / 2, 1e-300, 1e-9]])
    _, s, c = handle.math_probe(x)
    rs, rc = np.sin(x), np.cos(x)
    assert _ulps(s, rs).max() <= 2.0 and _ulps(c, rc).max() <= 2.0, (_ulps(s, rs).max(), _ulps(c, rc).max())
    assert np.abs(s - rs).max() < 2.3e-16 and np.abs(c - rc).max() < 2.3e-16
    # the values every equatorial ray lives on (theta = pi/2): identical to glibc's, so dH/dtheta's
    # cos(pi/2) = 6.1e-17 residual is reproduced bit-for-bit
    _, s, c = handle.math_probe(np.array([np.pi / 2]))
    assert s[0] == 1.0 and c[0] == np.cos(np.pi / 2) == 6.123233995736766e-17
    _, s, c = handle.math_probe(np.array([np.nan, np.inf]))
    assert np.isnan(s).all() and np.isnan(c).all()
