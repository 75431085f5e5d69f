"""The oracle's deterministic elementary functions against libm (numpy): <= 2 ulp (tan: quotient of two
1-ulp results), pow to 1e-13 relative (exp(y log x) without extended precision, see oracle/detmath.h)."""
import numpy as np


def _ulps(a, ref):
    return np.abs(a - ref) / np.spacing(np.abs(ref))


def test_sin_cos_tan(oracle_mod):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e3, 1e3, 200000), rng.uniform(-1e-3, 1e-3, 1000),
                        np.array([0.0, np.pi, -np.pi, np.pi / 2, 1e5, -1e5])])
    assert _ulps(oracle_mod.detmath(0, x), np.sin(x)).max() <= 1.0
    assert _ulps(oracle_mod.detmath(1, x), np.cos(x)).max() <= 1.0
    assert _ulps(oracle_mod.detmath(2, x), np.tan(x)).max() <= 2.5


def test_log_exp(oracle_mod):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(1e-300, 10, 200000), np.exp(rng.uniform(-700, 700, 200000)),
                        1.0 + rng.uniform(-0.3, 0.3, 100000), np.array([1.0, 5e-324, 2.2e-308])])
    assert _ulps(oracle_mod.detmath(3, x), np.log(x)).max() <= 1.0
    y = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-1, 1, 100000)])
    assert _ulps(oracle_mod.detmath(4, y), np.exp(y)).max() <= 1.0
    sp = oracle_mod.detmath(3, np.array([0.0, -1.0, np.inf, np.nan]))
    assert sp[0] == -np.inf and np.isnan(sp[1]) and sp[2] == np.inf and np.isnan(sp[3])
    assert oracle_mod.detmath(4, np.array([800.0]))[0] == np.inf
    assert oracle_mod.detmath(4, np.array([-800.0]))[0] == 0.0


def test_pow(oracle_mod):
    rng = np.random.default_rng(2)
    x = rng.uniform(1e-9, 3, 100000)
    y = rng.uniform(-3, 4, 100000)
    r = oracle_mod.detmath(5, x, y)
    assert np.max(np.abs(r - x ** y) / np.abs(x ** y)) < 1e-13
    assert oracle_mod.detmath(5, np.array([0.0]), np.array([1.1]))[0] == 0.0
    assert oracle_mod.detmath(5, np.array([2.0]), np.array([0.0]))[0] == 1.0
