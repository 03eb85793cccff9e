"""`chex` names used as annotations / dataclass decorator by the reference."""
import dataclasses

import torch
from jax.random import Key as PRNGKey  # noqa: F401

Array = torch.Tensor
ArrayDevice = torch.Tensor


def dataclass(cls=None, **kw):
    return dataclasses.dataclass(cls) if cls is not None else dataclasses.dataclass
