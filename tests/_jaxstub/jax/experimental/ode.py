"""[third-party, restated] `jax.experimental.ode.odeint`: adaptive Dormand-Prince 5(4).

The reference calls it at jax_ode.py:39-45, 100-106, 118-124 (JAX "v0.4.23" per the reference README.md:80; no lockfile).
Its source is not under /root/reference nor anywhere in this image, so the published algorithm is restated here:
Dormand-Prince tableau with FSAL, Hairer-Norsett-Wanner II.4 initial step (order 4), RMS error ratio over the WHOLE
flattened state against atol + rtol*max(|y0|,|y1|), step factor clip(0.9*ratio^(-1/5), 0.2 (1.0 when ratio < 1), 10)
(x10 when ratio == 0), and Shampine's 4th-order interpolant evaluated at every requested output time.

Differentiable by torch autograd through the accepted steps (step sizes are constants); the reference reaches the
gradient through the continuous adjoint `_odeint_rev` instead -- both converge to the exact gradient at O(tol).
"""
from __future__ import annotations

import torch

_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_BETA = [[1 / 5],
         [3 / 40, 9 / 40],
         [44 / 45, -56 / 15, 32 / 9],
         [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
         [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
         [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_C_ERR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
          -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.]
_C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
          187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]

last_stats = {}


def _lincomb(coeffs, ks):
    out = None
    for c, k in zip(coeffs, ks):
        if c != 0:
            out = c * k if out is None else out + c * k
    return out


def _initial_step_size(fun, t0, y0, order, rtol, atol, f0):
    scale = atol + y0.abs() * rtol
    d0 = torch.linalg.vector_norm((y0 / scale).reshape(-1)).item()
    d1 = torch.linalg.vector_norm((f0 / scale).reshape(-1)).item()
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = fun(y0 + h0 * f0, t0 + h0)
    d2 = torch.linalg.vector_norm(((f1 - f0) / scale).reshape(-1)).item() / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / (order + 1.0))
    return min(100.0 * h0, h1)


def _optimal_step_size(last_step, ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5.0):
    if ratio == 0:
        return last_step * ifactor
    dfactor = 1.0 if ratio < 1 else dfactor
    factor = min(ifactor, max(ratio ** (-1.0 / order) * safety, dfactor))
    return last_step * factor


def odeint(func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, mxstep=float("inf"), hmax=float("inf")):
    """ys[i] = y(t[i]); `func(y, t, *args)`.  One adaptive step sequence over all output times."""
    ts = [float(v) for v in t]
    f = lambda y, tt: func(y, torch.as_tensor(tt, dtype=torch.float64), *args)  # noqa: E731
    y = y0
    fy = f(y, ts[0])
    nfe = 1
    with torch.no_grad():
        dt = min(max(_initial_step_size(lambda yy, tt: f(yy.detach(), tt), ts[0], y.detach(), 4, rtol, atol, fy.detach()), 0.0), hmax)
    nfe += 1
    tcur = last_t = ts[0]
    coeff = [y0] * 5
    n_acc = n_rej = 0
    outs = [y0]
    for target in ts[1:]:
        while tcur < target and (n_acc + n_rej) < mxstep and dt > 0:
            ks = [fy]
            for i in range(6):
                yi = y + dt * _lincomb(_BETA[i], ks)
                ks.append(f(yi, tcur + dt * _ALPHA[i]))
                nfe += 1
            y1 = y + dt * _lincomb(_C_SOL, ks)
            with torch.no_grad():
                err = dt * _lincomb(_C_ERR, ks)
                tol = atol + rtol * torch.maximum(y.abs(), y1.abs())
                ratio = float(torch.sqrt(torch.mean((err / tol) ** 2)))
            new_dt = min(max(_optimal_step_size(dt, ratio), 0.0), hmax)
            if ratio <= 1.0:
                ymid = y + dt * _lincomb(_C_MID, ks)
                dy0, dy1 = ks[0], ks[6]
                a = -2. * dt * dy0 + 2. * dt * dy1 - 8. * y - 8. * y1 + 16. * ymid
                b = 5. * dt * dy0 - 3. * dt * dy1 + 18. * y + 14. * y1 - 32. * ymid
                c = -4. * dt * dy0 + dt * dy1 - 11. * y - 5. * y1 + 16. * ymid
                coeff = [a, b, c, dt * dy0, y]
                last_t, tcur, y, fy = tcur, tcur + dt, y1, ks[6]
                n_acc += 1
            else:
                n_rej += 1
            dt = new_dt
        rel = (target - last_t) / (tcur - last_t)
        out = coeff[0]
        for cc in coeff[1:]:
            out = out * rel + cc
        outs.append(out)
    last_stats.clear()
    last_stats.update(accepted=n_acc, rejected=n_rej, nfe=nfe)
    return torch.stack(outs, 0)
