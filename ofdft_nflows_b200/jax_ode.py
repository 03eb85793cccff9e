"""Host mirror of ofdft_normflows/jax_ode.py: ``neural_ode`` and ``neural_ode_score`` with the reference call
signatures ``(params, batch, f, t0, t1, d_dim)`` (jax_ode.py:9,51), backed by the fused fixed-step RK4 kernels.

Differences from the reference that a caller can observe:
  * the integrator is fixed-step RK4 (``n_steps`` knob, default ``DEFAULT_N_STEPS``) instead of adaptive Dopri5 at
    rtol=atol=1e-7 -- results agree to the discretisation error (DESIGN.md has measured deltas);
  * gradients w.r.t. ``params`` flow through torch.autograd (the analogue of jax.custom_vjp): the backward is the
    discrete adjoint kernel, not a continuous adjoint solve.
"""
from __future__ import annotations

from typing import Any, Tuple

import torch

from . import ops

DEFAULT_N_STEPS = 16


class _GnnOde(torch.autograd.Function):
    """y(t1) = RK4-flow(theta, y(t0)); backward = ofdft_cnf_gnn_bwd on the checkpointed stage states."""

    @staticmethod
    def forward(ctx, theta, batch, field, t0, t1, jets, n_steps):
        B = batch.shape[0]
        need_grad = theta.requires_grad or batch.requires_grad
        y1, ckpt = ops.gnn_fwd(theta, field.xyz_nuclei, field.z_one_hot, batch, B, jets, jets, n_steps,
                               float(t0), float(t1), field.y0, want_ckpt=need_grad)
        ctx.field, ctx.cfg = field, (float(t0), float(t1), jets, n_steps, B)
        ctx.save_for_backward(theta.detach(), ckpt if ckpt is not None else torch.empty(0, device=batch.device))
        ctx.want_ybar0 = batch.requires_grad
        return y1

    @staticmethod
    def backward(ctx, ybar1):
        theta, ckpt = ctx.saved_tensors
        t0, t1, jets, n_steps, B = ctx.cfg
        f = ctx.field
        tbar, ybar0 = ops.gnn_bwd(theta, f.xyz_nuclei, f.z_one_hot, ckpt, ybar1.contiguous(), B, jets, jets, n_steps,
                                  t0, t1, f.y0, want_ybar0=ctx.want_ybar0)
        return tbar, ybar0, None, None, None, None, None


def _integrate(params: Any, batch: torch.Tensor, f: Any, t0: float, t1: float, jets: int, n_steps: int) -> torch.Tensor:
    from .equiv_flows import Gen_EqvFlow
    from .cn_flows import Gen_CNFSimpleMLP, mlp_integrate
    theta = f.flat_theta(params)
    if isinstance(f, Gen_EqvFlow):
        return _GnnOde.apply(theta, batch.to(torch.float32).contiguous(), f, t0, t1, jets, n_steps)
    if isinstance(f, Gen_CNFSimpleMLP):
        return mlp_integrate(theta, batch.to(torch.float32).contiguous(), f, t0, t1, jets, n_steps)
    raise TypeError(f"neural_ode: unsupported vector field {type(f).__name__}")


def neural_ode(params: Any, batch: torch.Tensor, f: Any, t0: float, t1: float, d_dim: int,
               n_steps: int = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """jax_ode.py:9-49.  ``batch`` (B, d+1) rows [x, logp] -> (z_t1 (B,d), logp_t1 (B,1))."""
    y1 = _integrate(params, batch, f, t0, t1, ops.JETS_XLOGP, n_steps or DEFAULT_N_STEPS)
    return y1[:, :d_dim], y1[:, d_dim:d_dim + 1]


def neural_ode_score(params: Any, batch: torch.Tensor, f: Any, t0: float, t1: float, d_dim: int,
                     n_steps: int = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """jax_ode.py:51-110.  ``batch`` (B, 2d+1) rows [x, logp, score] -> (z_t1, logp_t1 (B,1), score_t1 (B,d))."""
    y1 = _integrate(params, batch, f, t0, t1, ops.JETS_FULL, n_steps or DEFAULT_N_STEPS)
    return y1[:, :d_dim], y1[:, d_dim:d_dim + 1], y1[:, d_dim + 1:2 * d_dim + 1]
