"""`jax.numpy` names used by the reference hot path, over torch float64 (see ../../README.md)."""
from __future__ import annotations

import math as _math

import numpy as _onp
import torch

pi = _math.pi
inf = _math.inf
newaxis = None
float64 = torch.float64
float32 = torch.float32
int32 = torch.int32
int64 = torch.int64
ndarray = torch.Tensor


def _dtype(dtype):
    if dtype is None:
        return None
    if dtype is int:
        return torch.int64
    if dtype is float:
        return torch.float64
    if dtype is bool:
        return torch.bool
    if isinstance(dtype, torch.dtype):
        return dtype
    return {"float64": torch.float64, "float32": torch.float32, "int64": torch.int64, "int32": torch.int32}[_onp.dtype(dtype).name]


def asarray(x, dtype=None):
    dt = _dtype(dtype)
    if isinstance(x, torch.Tensor):
        return x if dt is None else x.to(dt)
    if isinstance(x, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in x):
        return stack([asarray(e) for e in x]) if dt is None else stack([asarray(e) for e in x]).to(dt)
    a = _onp.asarray(x)
    if dt is None:
        if a.dtype.kind == "f":
            dt = torch.float64
        elif a.dtype.kind in "iu":
            dt = torch.int64
        elif a.dtype.kind == "b":
            dt = torch.bool
    return torch.as_tensor(a, dtype=dt)


array = asarray


def _f(x):
    """Operand of a floating-point function: tensors pass through, Python numbers become float64."""
    if isinstance(x, torch.Tensor):
        return x if x.is_floating_point() else x.to(torch.float64)
    return torch.as_tensor(x, dtype=torch.float64)


def _axis_kw(axis, keepdims):
    kw = {}
    if axis is not None:
        kw["dim"] = axis
    if keepdims:
        kw["keepdim"] = True
    return kw


def sum(a, axis=None, keepdims=False):  # noqa: A001
    return torch.sum(asarray(a), **_axis_kw(axis, keepdims))


def mean(a, axis=None, keepdims=False):
    return torch.mean(_f(a), **_axis_kw(axis, keepdims))


def sqrt(x): return torch.sqrt(_f(x))
def log(x): return torch.log(_f(x))
def log2(x): return torch.log2(_f(x))
def log1p(x): return torch.log1p(_f(x))
def exp(x): return torch.exp(_f(x))
def abs(x): return torch.abs(asarray(x))  # noqa: A001
def sin(x): return torch.sin(_f(x))
def cos(x): return torch.cos(_f(x))
def tanh(x): return torch.tanh(_f(x))
def arctan(x): return torch.arctan(_f(x))
def arcsinh(x): return torch.arcsinh(_f(x))
def arccos(x): return torch.arccos(_f(x))
def arctan2(y, x): return torch.arctan2(_f(y), _f(x))
def isnan(x): return torch.isnan(_f(x))
def isinf(x): return torch.isinf(_f(x))
def multiply(a, b): return asarray(a) * asarray(b)
def matmul(a, b): return torch.matmul(asarray(a), asarray(b))
def dot(a, b): return torch.matmul(asarray(a), asarray(b))
def vdot(a, b): return torch.dot(asarray(a).reshape(-1), asarray(b).reshape(-1))
def maximum(a, b): return torch.maximum(_f(a), _f(b))
def minimum(a, b): return torch.minimum(_f(a), _f(b))
def power(a, b): return _f(a) ** b


def round(x, decimals=0):  # noqa: A001
    return torch.round(_f(x), decimals=decimals)


def where(cond, x, y):
    x, y = asarray(x), asarray(y)
    if x.is_floating_point() != y.is_floating_point():
        x, y = _f(x), _f(y)
    return torch.where(cond, x, y)


def clip(a, a_min=None, a_max=None, min=None, max=None):  # noqa: A002
    lo = a_min if a_min is not None else min
    hi = a_max if a_max is not None else max
    return torch.clamp(asarray(a), min=lo, max=hi)


def einsum(subscripts, *operands):
    return torch.einsum(subscripts, *[asarray(o) for o in operands])


def ones(shape, dtype=None): return torch.ones(shape, dtype=_dtype(dtype) or torch.float64)
def zeros(shape, dtype=None): return torch.zeros(shape, dtype=_dtype(dtype) or torch.float64)
def ones_like(a): return torch.ones_like(asarray(a))
def zeros_like(a): return torch.zeros_like(asarray(a))


def linspace(start, stop, num=50):
    return torch.linspace(float(start), float(stop), int(num), dtype=torch.float64)


def arange(*args, dtype=None):
    return torch.arange(*args, dtype=_dtype(dtype))


def atleast_1d(x):
    x = asarray(x)
    return x.reshape(1) if x.dim() == 0 else x


def hstack(tup):
    arrs = [atleast_1d(a) for a in tup]
    if any(a.is_floating_point() for a in arrs):
        arrs = [_f(a) for a in arrs]
    return torch.cat(arrs, 0) if arrs[0].dim() == 1 else torch.cat(arrs, 1)


def vstack(tup): return torch.vstack([asarray(a) for a in tup])
def stack(arrays, axis=0): return torch.stack([asarray(a) for a in arrays], axis)
def concatenate(arrays, axis=0): return torch.cat([asarray(a) for a in arrays], axis)


def trace(a, offset=0, axis1=0, axis2=1):
    return torch.diagonal(a, offset=offset, dim1=axis1, dim2=axis2).sum(-1)


def squeeze(a, axis=None):
    return torch.squeeze(a) if axis is None else torch.squeeze(a, axis)


def expand_dims(a, axis): return torch.unsqueeze(asarray(a), axis)
def reshape(a, shape): return torch.reshape(asarray(a), shape if isinstance(shape, (tuple, list)) else (shape,))
def transpose(a, axes=None): return asarray(a).permute(*axes) if axes is not None else asarray(a).T
def unique(a): return torch.unique(asarray(a))


def trapz(y, x=None, dx=1.0, axis=-1):
    return torch.trapezoid(y, x, dim=axis) if x is not None else torch.trapezoid(y, dx=dx, dim=axis)


def polyval(p, x):
    out = p[0]
    for c in p[1:]:
        out = out * x + c
    return out


class _Linalg:
    @staticmethod
    def norm(x, ord=None, axis=None, keepdims=False):  # noqa: A002
        x = _f(x)
        if axis is None:
            return torch.linalg.vector_norm(x.reshape(-1), ord=2 if ord is None else ord)
        return torch.linalg.vector_norm(x, ord=2 if ord is None else ord, dim=axis, keepdim=keepdims)


linalg = _Linalg()
