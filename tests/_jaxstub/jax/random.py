"""`jax.random` stand-in: keys are seed tuples feeding NumPy generators.  NOT JAX's threefry stream -- base samples and
parameters are *inputs* of the path and are stored in every fixture."""
import numpy as _onp
import torch


class Key:
    def __init__(self, seed):
        self.seed = tuple(seed) if isinstance(seed, (tuple, list)) else (int(seed),)

    def rng(self):
        return _onp.random.default_rng(list(self.seed))

    def __repr__(self):
        return f"Key{self.seed}"


def PRNGKey(seed):
    return Key(seed)


key = PRNGKey


def split(key, num=2):
    return tuple(Key(key.seed + (i,)) for i in range(int(num)))


def _shape(shape):
    return (shape,) if isinstance(shape, int) else tuple(shape)


def normal(key, shape=(), dtype=None):
    return torch.as_tensor(key.rng().standard_normal(_shape(shape)), dtype=torch.float64)


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return torch.as_tensor(key.rng().uniform(minval, maxval, _shape(shape)), dtype=torch.float64)


def categorical_from_probs(key, probs, n):
    p = _onp.asarray(probs, dtype=_onp.float64)
    return torch.as_tensor(key.rng().choice(p.shape[-1], size=int(n), p=p / p.sum()), dtype=torch.int64)
