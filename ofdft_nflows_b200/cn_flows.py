"""Host mirror of ofdft_normflows/cn_flows.py: ``Gen_CNFSimpleMLP`` (the 1-D driver LiH.py imports it as ``CNF``).

g(x,t) = MLP([y0 t, x]) with tanh hidden layers and a zero-initialised last layer, exact trace
(cn_flows.py:117-182).  Only ``in_out_dim == 1`` (what LiH.py:64-65 uses, features (512,512,512)) is instantiated
in CUDA; the unused HyperNetwork / CNFRicky / FourierFeatures classes of the reference (cn_flows.py:18-114) are out
of scope.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from .equiv_flows import _init_dense_stack, _ravel, _unravel


class Gen_CNFSimpleMLP:
    """``Gen_CNFSimpleMLP(in_out_dim, features, bool_neg=False)`` (cn_flows.py:166-170)."""

    def __init__(self, in_out_dim: Any, features: Sequence[int], bool_neg: bool = False, device: str = "cuda"):
        if int(in_out_dim) != 1:
            raise NotImplementedError("Gen_CNFSimpleMLP: only in_out_dim == 1 is supported (LiH.py)")
        feats = tuple(int(f) for f in features)
        if len(set(feats)) != 1 or not (2 <= len(feats) <= 4) or feats[0] % 64 or not (64 <= feats[0] <= 512):
            raise NotImplementedError("Gen_CNFSimpleMLP: CUDA kernels need 2-4 equal hidden widths, multiple of 64, <= 512")
        self.in_out_dim, self.features = 1, feats
        self.H, self.n_layers = feats[0], len(feats)
        self.bool_neg = bool(bool_neg)
        self.y0 = -1.0 if self.bool_neg else 1.0
        self.device = torch.device(device)
        self.n_in = 2

    def init(self, key: Any = 0, t: Any = None, test_inputs: Any = None) -> Dict[str, Any]:
        return {"params": {"cnf": {"net": _init_dense_stack(int(key), 2, self.features, 1, self.device)}}}

    def apply(self, params: Dict[str, Any], t: Any, states: torch.Tensor) -> torch.Tensor:
        """``f.apply(params, t, states)`` (jax_ode.py:37, cn_flows.py:173-182): d/dt [x, logp] for states (B, 2); rows
        (B, 3) also get the score channel of ``_evol_fn_i`` (jax_ode.py:80-98)."""
        jets = {1: ops.JETS_X, 2: ops.JETS_XLOGP, 3: ops.JETS_FULL}[states.shape[1]]
        return mlp_rhs(self.flat_theta(params), states, self, float(t), jets)

    def flat_theta(self, params: Dict[str, Any]) -> torch.Tensor:
        return _ravel(params["params"]["cnf"]["net"])

    def unravel(self, flat: torch.Tensor) -> Dict[str, Any]:
        return {"params": {"cnf": {"net": _unravel(flat, 2, self.features, 1)}}}

    def theta_size(self) -> int:
        H, L = self.H, self.n_layers
        return 1 + H + (H + 2 * H) + (L - 1) * (H + H * H)


CNF = Gen_CNFSimpleMLP   # LiH.py alias


def mlp_rhs(theta, y, f: Gen_CNFSimpleMLP, t: float, jets: int) -> torch.Tensor:
    """ofdft_cnf_mlp1d_rhs: one evaluation of the vector field / score-augmented right-hand side at time t."""
    lib = _lib.load()
    theta, y = ops._f32c(theta), ops._f32c(y)
    B, stride = y.shape
    dy = torch.zeros_like(y)
    sb = lib.ofdft_cnf_mlp1d_scratch_bytes(B, f.H, f.n_layers, 1)
    scratch = torch.empty(sb, dtype=torch.uint8, device=y.device)
    with ops._on(y):
        _lib.check(lib.ofdft_cnf_mlp1d_rhs(ops._stream(y), ops._ptr(theta), ops._ptr(y), ops._ptr(dy), ops._ptr(scratch), sb, B,
                                           jets, stride, f.H, f.n_layers, float(t), f.y0), "ofdft_cnf_mlp1d_rhs")
    return dy


ACT_CKPT_MAX_BYTES = 16 << 30     # default cap of the 1-D activation checkpoint (~2.4 MB per row at 3 x 512, RK4 x 16)


def mlp_fwd(theta, batch, f: Gen_CNFSimpleMLP, t0: float, t1: float, jets: int, n_steps: int, want_ckpt: bool = False,
            path: int = ops.PATH_DEFAULT, act_ckpt: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """ofdft_cnf_mlp1d_fwd: rows [x | logp | s].  With ``want_ckpt`` the stage checkpoint the adjoint needs is returned; on
    the tensor-core path with full jets it also holds every evaluation's layer activations (``act_ckpt``, when the shape
    supports it and it fits ACT_CKPT_MAX_BYTES), which the adjoint then reads instead of recomputing the hidden layers --
    mlp_bwd recognises the larger buffer by its size."""
    lib = _lib.load()
    theta, batch = ops._f32c(theta), ops._f32c(batch)
    B, stride = batch.shape
    y1 = torch.zeros_like(batch)
    ckpt = None
    if want_ckpt:
        n = lib.ofdft_cnf_mlp1d_ckpt_bytes(B, n_steps)
        if act_ckpt and jets == ops.JETS_FULL and not (path & (ops.PATH_FFMA | ops.PATH_FFMA_DW)):
            n_act = lib.ofdft_cnf_mlp1d_act_ckpt_bytes(B, f.H, f.n_layers, n_steps)
            if 0 < n_act <= ACT_CKPT_MAX_BYTES:
                n, path = n_act, path | ops.PATH_ACT_CKPT
        ckpt = torch.empty(n // 4, dtype=torch.float32, device=batch.device)
    sb = lib.ofdft_cnf_mlp1d_scratch_bytes(B, f.H, f.n_layers, n_steps)
    scratch = torch.empty(sb, dtype=torch.uint8, device=batch.device)
    with ops._on(batch):
        _lib.check(lib.ofdft_cnf_mlp1d_fwd(ops._stream(batch), ops._ptr(theta), ops._ptr(batch), ops._ptr(y1), ops._ptr(ckpt),
                                           ops._ptr(scratch), sb, B, jets, stride, f.H, f.n_layers, n_steps,
                                           float(t0), float(t1), f.y0, int(path)), "ofdft_cnf_mlp1d_fwd")
    return y1, ckpt


def mlp_bwd(theta, ckpt, ybar1, f: Gen_CNFSimpleMLP, t0: float, t1: float, jets: int, n_steps: int,
            want_ybar0: bool = False, path: int = ops.PATH_DEFAULT):
    """ofdft_cnf_mlp1d_bwd: discrete adjoint; returns (theta_bar, ybar0 or None)."""
    lib = _lib.load()
    theta, ybar1 = ops._f32c(theta), ops._f32c(ybar1)
    B, stride = ybar1.shape
    tbar = torch.empty_like(theta)
    ybar0 = torch.zeros_like(ybar1) if want_ybar0 else None
    if B > 0 and ckpt.numel() * 4 > lib.ofdft_cnf_mlp1d_ckpt_bytes(B, n_steps):       # forward kept the layer activations
        if ckpt.numel() * 4 != lib.ofdft_cnf_mlp1d_act_ckpt_bytes(B, f.H, f.n_layers, n_steps) or (path & (ops.PATH_FFMA | ops.PATH_FFMA_DW)):
            raise ValueError("mlp_bwd: checkpoint size matches neither the stage nor the activation checkpoint of this shape/path")
        path |= ops.PATH_ACT_CKPT
    sb = lib.ofdft_cnf_mlp1d_scratch_bytes(B, f.H, f.n_layers, n_steps)
    scratch = torch.empty(sb, dtype=torch.uint8, device=ybar1.device)
    with ops._on(ybar1):
        _lib.check(lib.ofdft_cnf_mlp1d_bwd(ops._stream(ybar1), ops._ptr(theta), ops._ptr(ckpt), ops._ptr(ybar1), ops._ptr(tbar),
                                           ops._ptr(ybar0), ops._ptr(scratch), sb, B, jets, stride, f.H, f.n_layers, n_steps,
                                           float(t0), float(t1), f.y0, int(path)), "ofdft_cnf_mlp1d_bwd")
    return tbar, ybar0


class _MlpOde(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, batch, field, t0, t1, jets, n_steps):
        need = theta.requires_grad or batch.requires_grad
        if need and jets != ops.JETS_FULL:
            # the adjoint kernels carry all three channels: differentiate neural_ode through neural_ode_score instead
            # (pad the rows with a zero score column and leave its cotangent zero)
            raise NotImplementedError("Gen_CNFSimpleMLP: gradients are implemented for neural_ode_score (rows [x, logp, score]); "
                                      "pad the batch with a score column to differentiate neural_ode")
        y1, ckpt = mlp_fwd(theta, batch, field, t0, t1, jets, n_steps, want_ckpt=need)
        ctx.field, ctx.cfg = field, (float(t0), float(t1), jets, n_steps)
        ctx.save_for_backward(theta.detach(), ckpt if ckpt is not None else torch.empty(0, device=batch.device))
        ctx.want_ybar0 = batch.requires_grad
        return y1

    @staticmethod
    def backward(ctx, ybar1):
        theta, ckpt = ctx.saved_tensors
        t0, t1, jets, n_steps = ctx.cfg
        tbar, ybar0 = mlp_bwd(theta, ckpt, ybar1.contiguous(), ctx.field, t0, t1, jets, n_steps, want_ybar0=ctx.want_ybar0)
        return tbar, ybar0, None, None, None, None, None


def mlp_integrate(theta, batch, f, t0, t1, jets, n_steps):
    return _MlpOde.apply(theta, batch, f, t0, t1, jets, n_steps)
