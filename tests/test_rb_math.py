"""CPU: csrc/rb_math.h -- the sin / cos / exp the CUDA kernels and the oracle share -- against libm.

The routines exist to make GPU and oracle agree bit for bit; this test pins how far they are from the
libm the reference's NumPy / math calls use: never more than 1 ulp, and equal in all but a few per mille
of the arguments (where they are the correctly rounded ones more often than not)."""
import numpy as np

import oracle as O


def ulps(a, b):
    return np.abs(a.view(np.int64) - b.view(np.int64))


def test_sincos_within_one_ulp_of_libm():
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.uniform(-4, 4, 200000), rng.uniform(-100, 100, 200000), rng.uniform(-1e5, 1e5, 100000),
                        np.array([0.0, np.pi / 2, -np.pi / 2, np.pi, 1e-300, -1e-300, 0.7853981633974483])])
    s, c = O.rb_sincos(a)
    us, uc = ulps(s, np.sin(a)), ulps(c, np.cos(a))
    assert us.max() <= 1 and uc.max() <= 1
    assert np.mean(us > 0) < 5e-3 and np.mean(uc > 0) < 5e-3
    assert s[-7] == 0.0 and c[-7] == 1.0
    sn, cn = O.rb_sincos(np.array([np.inf, np.nan]))
    assert np.isnan(sn).all() and np.isnan(cn).all()


def test_exp_within_one_ulp_of_libm():
    rng = np.random.default_rng(2)
    import math

    x = np.concatenate([-rng.uniform(0, 40, 60000), -rng.uniform(0, 700, 40000), rng.uniform(0, 700, 10000),
                        np.array([0.0, 1.0, -745.0, -750.0, -1e9, 709.7])])
    e = O.rb_exp(x)
    ref = np.array([math.exp(v) for v in x])          # glibc (np.exp is NumPy's own SIMD routine, 1-2 ulp off glibc in 5 % of arguments)
    u = ulps(e, ref)
    assert u.max() <= 1
    assert np.mean(u > 0) < 5e-3
    assert e[-6] == 1.0 and e[-3] == 0.0 and e[-2] == 0.0
    assert O.rb_exp(np.array([800.0]))[0] == np.inf and np.isnan(O.rb_exp(np.array([np.nan]))[0])
