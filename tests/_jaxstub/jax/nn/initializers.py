"""`jax.nn.initializers`: (key, shape, dtype) -> array.  Random draws come from the stub key's NumPy generator (JAX's
PRNG is not reproduced; parameters are inputs of every fixture)."""
import math

import torch


def zeros(key, shape, dtype=torch.float64):
    return torch.zeros(tuple(shape), dtype=torch.float64)


def ones(key, shape, dtype=torch.float64):
    return torch.ones(tuple(shape), dtype=torch.float64)


def normal(stddev=1e-2, dtype=torch.float64):
    def init(key, shape, dtype=dtype):
        return torch.as_tensor(key.rng().standard_normal(tuple(shape)) * stddev, dtype=torch.float64)
    return init


def lecun_normal(dtype=torch.float64):
    """variance_scaling(1.0, 'fan_in', 'truncated_normal'): N(0, 1/fan_in) truncated at +-2 sigma and rescaled."""
    def init(key, shape, dtype=dtype):
        fan_in = shape[0]
        std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
        g = key.rng()
        x = g.standard_normal(tuple(shape))
        bad = abs(x) > 2.0
        while bad.any():
            x[bad] = g.standard_normal(int(bad.sum()))
            bad = abs(x) > 2.0
        return torch.as_tensor(x * std, dtype=torch.float64)
    return init
