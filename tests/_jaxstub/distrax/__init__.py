"""[third-party, restated] the `distrax` classes the reference priors use (promolecular_distrax.py:18-89, LiH.py:78).
Sampling uses NumPy generators (not JAX's PRNG): samples are inputs of the path."""
import math

import torch
from jax import random as _jr


class Distribution:
    def sample(self, *, seed, sample_shape=()):
        n = int(sample_shape) if not isinstance(sample_shape, (tuple, list)) else int(torch.tensor(sample_shape).prod())
        out = self._sample_n(seed, n)
        if isinstance(sample_shape, (tuple, list)):
            out = out.reshape(tuple(sample_shape) + tuple(out.shape[1:]))
        return out

    def log_prob(self, value):
        raise NotImplementedError

    def prob(self, value):
        return torch.exp(self.log_prob(value))


class MultivariateNormalDiag(Distribution):
    """Independent normal components over the LAST axis; leading axes of loc / scale_diag are batch axes."""

    def __init__(self, loc=None, scale_diag=None):
        self.loc = torch.as_tensor(loc, dtype=torch.float64)
        self.scale_diag = torch.ones_like(self.loc) if scale_diag is None else torch.as_tensor(scale_diag, dtype=torch.float64)

    def log_prob(self, value):
        z = (value - self.loc) / self.scale_diag
        return torch.sum(-0.5 * z * z - torch.log(self.scale_diag) - 0.5 * math.log(2.0 * math.pi), -1)

    def _sample_n(self, key, n):
        shape = (int(n),) + tuple(torch.broadcast_shapes(self.loc.shape, self.scale_diag.shape))
        return self.loc + self.scale_diag * _jr.normal(key, shape)


class Categorical(Distribution):
    def __init__(self, logits=None, probs=None):
        if probs is None:
            probs = torch.softmax(torch.as_tensor(logits, dtype=torch.float64), -1)
        p = torch.as_tensor(probs, dtype=torch.float64)
        self.probs = p / p.sum(-1, keepdim=True)

    def _sample_n(self, key, n):
        return _jr.categorical_from_probs(key, self.probs.numpy(), n)


class MixtureSameFamily(Distribution):
    def __init__(self, mixture_distribution, components_distribution):
        self.mixture_distribution, self.components_distribution = mixture_distribution, components_distribution
