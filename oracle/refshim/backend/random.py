"""The JAX PRNG is not reproduced; only what the deterministic step loop might import."""
import numpy as _np


def prng_key(*, seed):
    return _np.random.Generator(_np.random.PCG64(seed))


def split(key, num):
    """`num` independent generators, as an ARRAY of generators (JAX returns an array of keys, and the reference maps
    over it with vmap in `MarkovSequence.sample`)."""
    out = _np.empty((num,), dtype=object)
    for i in range(num):
        out[i] = _np.random.Generator(_np.random.PCG64(int(key.integers(0, 2**31))))
    return out


def normal(key, /, shape, dtype=None):
    return key.normal(size=shape).astype(dtype or _np.float64)


def rademacher(key, /, shape, dtype):
    return (2 * key.integers(0, 2, size=shape) - 1).astype(dtype)


def logpdf_multivariate_normal(x, /, mean, cov):
    import scipy.stats

    return scipy.stats.multivariate_normal.logpdf(x, mean, cov)
