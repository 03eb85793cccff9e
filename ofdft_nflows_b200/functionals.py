"""Host mirror of the string-keyed functional factories of ofdft_normflows/functionals.py:
``_kinetic`` (:21-43), ``_exchange_correlation`` (:164-183), ``_hartree`` (:425-436), ``_nuclear`` (:504-521).

Each factory returns a callable with the reference's argument order and ``(bs, 1)`` result; the arithmetic and the
derivatives run in ``ofdft_functional_eval`` (CUDA), and ``torch.autograd`` plays the role of the JAX VJP so that a
``loss`` written like the reference's (H20_mol_ofdft_min_equiv_flows.py:150-173) differentiates end to end.  The
fused ``EnergyEstimator`` step does not go through these wrappers (one kernel evaluates all terms and their
cotangents); they exist for drop-in use and for per-term inspection.

Covered: every functional the reference drivers select by default (tf / w / tf-w / tf1d / tfw_1d; hartree / softc /
hartree_mt; nuclei_potential / hgh / attraction; lda / b88_x_e / dirac_b88_x_e; vwn_c_e / pw92_c_e / xc_1d).  The unused
``kinetic`` / ``madness`` / ``harmonic`` variants raise NotImplementedError.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Dict

import numpy as np
import torch

from . import _lib, ops
from .ops import EnergySpec


def _eval(family: int, spec: EnergySpec, a0: torch.Tensor, a1, want_grad: bool):
    lib = _lib.load()
    a0 = ops._f32c(a0)
    a1 = ops._f32c(a1) if a1 is not None else None
    n = a0.shape[0]
    cs, coords, z = spec.to_c()
    out = torch.empty((n, 1), dtype=torch.float32, device=a0.device)
    g0 = torch.empty_like(a0) if want_grad else None
    g1 = torch.empty_like(a1) if (want_grad and a1 is not None) else None
    with ops._on(a0):
        _lib.check(lib.ofdft_functional_eval(ops._stream(a0), family, C.byref(cs), ops._host_ptr(coords), ops._host_ptr(z),
                                             ops._ptr(a0), ops._ptr(a1), n, ops._ptr(out), ops._ptr(g0), ops._ptr(g1)),
                   "ofdft_functional_eval")
    return out, g0, g1


class _Functional(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a0, a1, family, spec):
        need = a0.requires_grad or (a1 is not None and a1.requires_grad)
        out, g0, g1 = _eval(family, spec, a0, a1, need)
        ctx.save_for_backward(*(t for t in (g0, g1) if t is not None))
        ctx.has = (g0 is not None, g1 is not None)
        return out

    @staticmethod
    def backward(ctx, gout):
        saved = list(ctx.saved_tensors)
        g0 = saved.pop(0) if ctx.has[0] else None
        g1 = saved.pop(0) if ctx.has[1] else None
        n = gout.shape[0]
        mul = lambda g: None if g is None else (gout.reshape(n, 1) * g.reshape(n, -1)).reshape(g.shape)
        return mul(g0), mul(g1), None, None


def _kinetic(name: str = "TF") -> Callable:
    """functionals.py:21-43: ``wrapper(den, score, Ne) -> (bs,1)``."""
    key = name.lower()
    if key not in ops.KIN or ops.KIN[key] == 0:
        raise NotImplementedError(f"_kinetic({name!r}) is not implemented by the CUDA path")

    def wrapper(den, score, Ne):
        d = score.shape[1]
        return _Functional.apply(den.reshape(-1), score, 0, EnergySpec(Ne=float(Ne), d=d, kin=key, hart="", nuc="", x=""))
    return wrapper


def _exchange_correlation(name: str = "XC") -> Callable:
    """functionals.py:164-183.  x-type (``lda``, ``b88_x_e``, ``dirac_b88_x_e``): ``wrapper(den, score, Ne)``;
    c-type (``vwn_c_e``, ``pw92_c_e``, ``xc_1d``): ``wrapper(den, Ne)``."""
    key = name.lower()
    if key in ops.XC_C and ops.XC_C[key] != 0:
        def wrapper_c(den, Ne):
            d = 1 if key in ("xc_1d", "exchange_correlation_1d") else 3
            return _Functional.apply(den.reshape(-1), None, 4, EnergySpec(Ne=float(Ne), d=d, kin="", hart="", nuc="", x="", c=key))
        return wrapper_c
    if key not in ops.XC_X or ops.XC_X[key] == 0:
        raise NotImplementedError(f"_exchange_correlation({name!r}) is not implemented by the CUDA path")

    def wrapper(den, score, Ne):
        d = score.shape[1]
        return _Functional.apply(den.reshape(-1), score, 1, EnergySpec(Ne=float(Ne), d=d, kin="", hart="", nuc="", x=key))
    return wrapper


def _hartree(name: str = "HP") -> Callable:
    """functionals.py:425-436: ``wrapper(x, xp, Ne) -> (bs,1)`` (row-paired estimator)."""
    key = name.lower()
    if key not in ops.HART or ops.HART[key] == 0:
        raise NotImplementedError(f"_hartree({name!r}) is not implemented by the CUDA path")

    def wrapper(x, xp, Ne):
        d = x.shape[1]
        return _Functional.apply(x, xp, 2, EnergySpec(Ne=float(Ne), d=d, kin="", hart=key, nuc="", x=""))
    return wrapper


def _nuclear(name: str = "NP") -> Callable:
    """functionals.py:504-521: 3-D ``wrapper(x, Ne, mol_info)`` with ``mol_info = {'coords': (Na,3), 'z': (Na,)}``
    (H20_...py:76); 1-D ``attraction``: ``wrapper(x, R, Z_alpha, Z_beta, Ne)`` (functionals.py:523-549)."""
    key = name.lower()
    if key in ("attraction", "attr"):
        def wrapper(x, R, Z_alpha, Z_beta, Ne):
            spec = EnergySpec(Ne=float(Ne), d=1, kin="", hart="", nuc=key, x="", R=float(R), Z_alpha=float(Z_alpha), Z_beta=float(Z_beta))
            return _Functional.apply(x, None, 3, spec)
        return wrapper
    if key not in ops.NUC or ops.NUC[key] == 0:
        raise NotImplementedError(f"_nuclear({name!r}) is not implemented by the CUDA path")

    def wrapper3(x, Ne, mol_info: Dict[str, Any]):
        coords = np.asarray(mol_info["coords"].detach().cpu() if isinstance(mol_info["coords"], torch.Tensor) else mol_info["coords"])
        z = np.asarray(mol_info["z"].detach().cpu() if isinstance(mol_info["z"], torch.Tensor) else mol_info["z"])
        spec = EnergySpec(Ne=float(Ne), d=3, kin="", hart="", nuc=key, x="", coords=coords, z=z)
        return _Functional.apply(x, None, 3, spec)
    return wrapper3
