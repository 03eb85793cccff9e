"""Test-only stand-in for the `jax` names the reference's hot-path files use (see ../README.md).

Backed by torch (CPU, float64) and torch.func.  NOT a JAX implementation: only what
/root/reference/ofdft_normflows/{functionals,utils,jax_ode,equiv_flows,cn_flows,promolecular_distrax,
hgh_pseudopotentials}.py touch, with JAX's semantics for those calls (x64 enabled, as every reference driver sets
`jax_enable_x64`, e.g. H20_mol_ofdft_min_equiv_flows.py:31).
"""
from __future__ import annotations

import torch
from torch import func as _tf

__version__ = "0.4.23+torchstub"

Array = torch.Tensor


class _Config:
    def __init__(self):
        self.values = {"jax_enable_x64": True}

    def update(self, key, value):
        self.values[key] = value


config = _Config()


def _t(x):
    """Python scalars -> float64 / int64 tensors (what jnp.asarray does under x64)."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, bool):
        return torch.tensor(x)
    if isinstance(x, int):
        return torch.tensor(x, dtype=torch.int64)
    if isinstance(x, float):
        return torch.tensor(x, dtype=torch.float64)
    return x


def jit(fun=None, static_argnums=None, static_argnames=None, **_):
    """Tracing is a no-op here: the function runs eagerly."""
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    return _tf.vmap(fun, in_dims=in_axes, out_dims=out_axes)


def _argwrap(f):
    def g(*args, **kw):
        return f(*[_t(a) for a in args], **kw)
    return g


def grad(fun, argnums=0, has_aux=False):
    return _argwrap(_tf.grad(fun, argnums=argnums, has_aux=has_aux))


def value_and_grad(fun, argnums=0, has_aux=False):
    gv = _tf.grad_and_value(fun, argnums=argnums, has_aux=has_aux)

    def wrapped(*args, **kw):
        g, v = gv(*[_t(a) for a in args], **kw)
        return v, g
    return wrapped


def vjp(fun, *primals, has_aux=False):
    return _tf.vjp(fun, *primals, has_aux=has_aux)


def jacrev(fun, argnums=0):
    return _argwrap(_tf.jacrev(fun, argnums=argnums))


def jacfwd(fun, argnums=0):
    return _argwrap(_tf.jacfwd(fun, argnums=argnums))


def hessian(fun, argnums=0):
    return _argwrap(_tf.hessian(fun, argnums=argnums))


class custom_jvp:
    """`@jax.custom_jvp` with `.defjvp`: primal evaluation only (equiv_flows.py:14-26 defines `safe_sqrt` this way and
    never calls it)."""

    def __init__(self, fun):
        self.fun = fun
        self.jvp = None

    def defjvp(self, jvp):
        self.jvp = jvp
        return jvp

    def __call__(self, *args):
        return self.fun(*args)


def device_put(x, *_, **__):
    return x


def device_get(x):
    return x


from . import numpy, lax, nn, random, tree_util, flatten_util  # noqa: E402,F401
from . import experimental, scipy, _src  # noqa: E402,F401
