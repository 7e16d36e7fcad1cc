"""The JAX PRNG is not reproduced; only what the deterministic step loop might import."""
import numpy as _np


def prng_key(*, seed):
    return _np.random.Generator(_np.random.PCG64(seed))


def split(key, num):
    return [_np.random.Generator(_np.random.PCG64(int(key.integers(0, 2**31)))) for _ in range(num)]


def normal(key, /, shape, dtype=None):
    return key.normal(size=shape).astype(dtype or _np.float64)


def rademacher(key, /, shape, dtype):
    return (2 * key.integers(0, 2, size=shape) - 1).astype(dtype)


def logpdf_multivariate_normal(x, /, mean, cov):
    import scipy.stats

    return scipy.stats.multivariate_normal.logpdf(x, mean, cov)
