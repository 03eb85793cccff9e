"""`jax.nn` names used by the reference hot path (see ../../README.md)."""
import torch

from ..numpy import asarray
from . import initializers  # noqa: F401


def one_hot(x, num_classes, dtype=None, axis=-1):
    """jax.nn.one_hot: `x[..., None] == arange(num_classes)` -- values that match no class (e.g. the FLOAT atomic numbers
    8. or 3. with num_classes = 3, H20_mol_ofdft_min_equiv_flows.py:83-84) give an all-zero row."""
    x = asarray(x)
    classes = torch.arange(int(num_classes), dtype=x.dtype)
    return (x[..., None] == classes).to(torch.float64 if dtype is None else dtype)


def tanh(x): return torch.tanh(x)
def sigmoid(x): return torch.sigmoid(x)
def softplus(x): return torch.nn.functional.softplus(x)
def relu(x): return torch.relu(x)
