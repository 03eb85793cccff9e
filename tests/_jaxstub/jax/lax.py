"""`jax.lax` names used by the reference hot path (see ../README.md)."""
import torch

from .numpy import asarray, _f


def expand_dims(array, dimensions):
    x = asarray(array)
    for d in sorted(int(d) for d in (dimensions if isinstance(dimensions, (tuple, list)) else (dimensions,))):
        x = x.unsqueeze(d)
    return x


def concatenate(operands, dimension):
    return torch.cat([asarray(o) for o in operands], int(dimension))


def reshape(operand, new_sizes, dimensions=None):
    return torch.reshape(operand, tuple(new_sizes))


def erf(x): return torch.erf(_f(x))
def erfc(x): return torch.erfc(_f(x))
def stop_gradient(x): return x.detach()
