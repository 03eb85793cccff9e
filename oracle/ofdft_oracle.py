"""CPU oracle for the ofdft_nflows Monte-Carlo energy-estimator hot path (NumPy float64).

TEST INFRASTRUCTURE ONLY.  Nothing under ``ofdft_nflows_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and only as the checker / the reported CPU baseline.

PARITY UNPINNED.  The reference (``/root/reference``, pure JAX/Flax) ships no tests, golden
vectors or fixtures for this path, and JAX/Flax/Distrax are not installable in the build image,
so the reference itself cannot be executed to pin this restatement.  What pins it instead:
  * ``oracle/ref_torch.py`` -- an independent float64 *autodiff transliteration* of the
    reference source (``vmap(jacrev)`` trace, ``value_and_grad`` + ``vjp`` score dynamics,
    nested reverse-mode for the parameter gradient); ``tests/test_oracle_*.py`` check the closed
    forms below against it;
  * size-independent properties (zero-init identity flow, rev o fwd = id, score = finite
    difference of log rho, closed-form functional values, N vs 2N self-convergence).

Every function cites the reference file:line it restates.  Third-party arithmetic that is NOT in
/root/reference (``jax.experimental.ode.odeint``: Dopri5 tableau, step controller, interpolant;
JAX "v0.4.23" per reference README.md:80, unpinned in SETUP.py:12) is restated from its
published algorithm (Dormand-Prince 5(4), Hairer-Norsett-Wanner II.4 initial step, Shampine
4th-order interpolant) and marked [third-party].

Flat parameter layout (``theta``), float, key-sorted like ``jax.flatten_util.ravel_pytree`` of the
Flax pytree:
  GNN  (equiv_flows.py:36-57):  last_layer.bias(1), last_layer.kernel(H_L,1),
        layers_0.bias(H_1), layers_0.kernel(I,H_1), layers_1.bias(H_2), layers_1.kernel(H_1,H_2), ...
        with I = 2 + n_types  (input order [t, r, onehot], equiv_flows.py:51)
  MLP  (cn_flows.py:117-141):   same ordering with I = 1 + d  (input order [t, x], cn_flows.py:133-134)
Kernels are (in, out) row-major as in flax.linen.Dense.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Sequence, Tuple

import numpy as np

BOHR = 1.8897259886  # H20_mol_ofdft_min_equiv_flows.py:32


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------

@dataclass
class MLPParams:
    """Dense stack: hidden[(W (in,out), b (out,)), ...] with tanh, then last (W (H,out), b (out,))."""
    hidden: List[Tuple[np.ndarray, np.ndarray]]
    last: Tuple[np.ndarray, np.ndarray]

    @property
    def n_in(self) -> int:
        return self.hidden[0][0].shape[0]

    def copy(self) -> "MLPParams":
        return MLPParams([(w.copy(), b.copy()) for w, b in self.hidden],
                         (self.last[0].copy(), self.last[1].copy()))


def theta_size(n_in: int, features: Sequence[int], n_out: int = 1) -> int:
    n = n_out + features[-1] * n_out
    prev = n_in
    for h in features:
        n += h + prev * h
        prev = h
    return n


def flatten_params(p: MLPParams) -> np.ndarray:
    """ravel_pytree order: last_layer.{bias,kernel}, layers_i.{bias,kernel}."""
    parts = [p.last[1].ravel(), p.last[0].ravel()]
    for w, b in p.hidden:
        parts += [b.ravel(), w.ravel()]
    return np.concatenate(parts)


def unflatten_params(theta: np.ndarray, n_in: int, features: Sequence[int], n_out: int = 1) -> MLPParams:
    theta = np.asarray(theta)
    o = 0
    hl = features[-1]
    b_last = theta[o:o + n_out]; o += n_out
    w_last = theta[o:o + hl * n_out].reshape(hl, n_out); o += hl * n_out
    hidden = []
    prev = n_in
    for h in features:
        b = theta[o:o + h]; o += h
        w = theta[o:o + prev * h].reshape(prev, h); o += prev * h
        hidden.append((w, b))
        prev = h
    assert o == theta.size, (o, theta.size)
    return MLPParams(hidden, (w_last, b_last))


def init_params(rng: np.random.Generator, n_in: int, features: Sequence[int], n_out: int = 1,
                last_scale: float = 0.0) -> MLPParams:
    """Flax Dense defaults: LeCun-normal kernels (std = 1/sqrt(fan_in)), zero bias; last layer
    zero-init (equiv_flows.py:43-45, cn_flows.py:124-126).  ``last_scale`` > 0 draws the last kernel
    as last_scale * N(0, 1/H) (benchmark/parity parameters must have a non-zero field).
    (flax uses a *truncated* normal; the distribution detail is irrelevant to parity: theta is an input.)"""
    hidden = []
    prev = n_in
    for h in features:
        hidden.append((rng.standard_normal((prev, h)) / math.sqrt(prev), np.zeros(h)))
        prev = h
    w_last = last_scale * rng.standard_normal((prev, n_out)) / math.sqrt(prev)
    return MLPParams(hidden, (w_last, np.zeros(n_out)))


# ----------------------------------------------------------------------------------------------
# 2-jet of the tanh MLP in one scalar input
# ----------------------------------------------------------------------------------------------

def mlp_jet(p: MLPParams, pre0: np.ndarray, dpre0: np.ndarray, n_jets: int = 3, keep: bool = False):
    """Propagate value / first / second derivative (w.r.t. one scalar input u) through the stack.

    pre0  : (..., H1) first-layer pre-activation   p  = W1^T [inputs] + b1
    dpre0 : (H1,)     its derivative in u          p' = W1[row of u, :]   (p'' = 0)
    Returns f, f', f'' of shape (..., n_out) (entries beyond n_jets are None) and, if ``keep``,
    the per-layer cache needed by :func:`mlp_jet_vjp`.
    tanh 2-jet: a = tanh p, a' = (1-a^2) p', a'' = (1-a^2) p'' - 2 a a' p'.
    """
    cache = []
    pz, p1, p2 = pre0, np.broadcast_to(dpre0, pre0.shape), None
    a = a1 = a2 = None
    for li, (w, b) in enumerate(p.hidden):
        if li > 0:
            pz = a @ w + b
            p1 = a1 @ w if n_jets >= 2 else None
            p2 = a2 @ w if n_jets >= 3 else None
        a_in = (a, a1, a2)
        a = np.tanh(pz)
        sg = 1.0 - a * a
        a1 = sg * p1 if n_jets >= 2 else None
        if n_jets >= 3:
            a2 = -2.0 * a * a1 * p1 if p2 is None else sg * p2 - 2.0 * a * a1 * p1
        if keep:
            cache.append((a_in, p1, p2, a, sg, a1, a2))
    wl, bl = p.last
    f = a @ wl + bl
    f1 = a1 @ wl if n_jets >= 2 else None
    f2 = a2 @ wl if n_jets >= 3 else None
    if keep:
        return f, f1, f2, cache
    return f, f1, f2


def mlp_jet_vjp(p: MLPParams, cache, fb, f1b, f2b):
    """Reverse pass of :func:`mlp_jet`.  fb/f1b/f2b: cotangents of f, f', f'' with shape (N, n_out)
    (None = zero).  Returns (pre0_bar (N,H1), dpre0_bar (H1,), grads) where grads mirrors MLPParams
    for every layer except the first kernel/bias (the caller owns the input map of layer 0).
    This is the closed-form counterpart of JAX reverse-mode through jacrev/grad (third derivatives
    of the network appear implicitly through a'' -> p)."""
    n = cache[0][3].shape[0]
    wl, bl = p.last
    a_in, p1, p2, a, sg, a1, a2 = cache[-1]
    zeros = lambda: np.zeros((n, wl.shape[0]))
    gw_last = np.zeros_like(wl)
    gb_last = np.zeros_like(bl)
    ab = zeros(); a1b = zeros(); a2b = zeros()
    if fb is not None:
        ab = fb @ wl.T; gw_last += a.T @ fb; gb_last += fb.sum(0)
    if f1b is not None and a1 is not None:
        a1b = f1b @ wl.T; gw_last += a1.T @ f1b
    if f2b is not None and a2 is not None:
        a2b = f2b @ wl.T; gw_last += a2.T @ f2b
    g_hidden = [None] * len(p.hidden)
    for li in range(len(p.hidden) - 1, -1, -1):
        a_in, p1, p2, a, sg, a1, a2 = cache[li]
        w, b = p.hidden[li]
        p1v = p1 if p1 is not None else 0.0
        p2v = p2 if p2 is not None else 0.0
        # a'' = sg p'' - 2 a sg p'^2 ; a' = sg p' ; a = tanh p
        p2b = a2b * sg
        p1b = sg * (a1b - 4.0 * a * p1v * a2b)
        pzb = ab * sg - 2.0 * a * sg * p1v * a1b + a2b * (-2.0 * a * sg * p2v - 2.0 * sg * (1.0 - 3.0 * a * a) * p1v * p1v)
        if li == 0:
            # first layer: p' is the constant row dpre0 (broadcast), p'' = 0
            return pzb, p1b.sum(0), (g_hidden, (gw_last, gb_last))
        ain, a1in, a2in = a_in
        gw = ain.T @ pzb
        if a1in is not None:
            gw += a1in.T @ p1b
        if a2in is not None:
            gw += a2in.T @ p2b
        g_hidden[li] = (gw, pzb.sum(0))
        ab = pzb @ w.T
        a1b = p1b @ w.T
        a2b = p2b @ w.T
    raise AssertionError


# ----------------------------------------------------------------------------------------------
# vector fields (closed form of equiv_flows.py / cn_flows.py + jax_ode.py:80-94)
# ----------------------------------------------------------------------------------------------

@dataclass
class GNNField:
    """Gen_EqvFlow (equiv_flows.py:108-127): g(x,t) = sum_i f_theta(t', |x-R_i|, onehot_i) (x-R_i),
    t' = y0*t, output scaled by y0 (= -1 if bool_neg).  State rows [x(3), logp, (score(3))]."""
    params: MLPParams
    nuclei: np.ndarray      # (Na, 3)
    onehot: np.ndarray      # (Na, n_types)
    bool_neg: bool = False
    d: int = 3

    @property
    def y0(self) -> float:
        return -1.0 if self.bool_neg else 1.0

    def _first_pre(self, tp: float, r: np.ndarray):
        """p = [t', r, onehot_i] @ W1 + b1 for all (row, nucleus): returns (N, Na, H), w_r (H,)."""
        w1, b1 = self.params.hidden[0]
        c = tp * w1[0] + self.onehot @ w1[2:] + b1            # (Na, H)
        return r[..., None] * w1[1] + c[None], w1[1]

    def rhs(self, t: float, y: np.ndarray, with_score: bool) -> np.ndarray:
        """d/dt of [x, logp] (jax_ode.py:36-37 via equiv_flows.py:96-105,124-127) or of
        [x, logp, score] (jax_ode.py:80-94):  dscore = -J^T s + grad_x(dlogp)."""
        y0 = self.y0
        x = y[:, :3]
        dvec = x[:, None, :] - self.nuclei[None]              # (N, Na, 3)
        r = np.sqrt((dvec * dvec).sum(-1))                    # jnp.linalg.norm, no eps (equiv_flows.py:77)
        pre, wr = self._first_pre(y0 * t, r)
        f, f1, f2 = mlp_jet(self.params, pre, wr, 3 if with_score else 2)
        f = f[..., 0]; f1 = f1[..., 0]
        g = (f[..., None] * dvec).sum(1)
        div = (3.0 * f + f1 * r).sum(1)
        out = [y0 * g, (-y0 * div)[:, None]]
        if with_score:
            f2 = f2[..., 0]
            s = y[:, 4:7]
            u = dvec / r[..., None]
            q = (dvec * s[:, None, :]).sum(-1)
            amp = f1 * (q + 4.0) + f2 * r                    # J s = f s + f' q u ; grad div = (4 f' + f'' r) u
            ds = -y0 * (f.sum(1)[:, None] * s + (amp[..., None] * u).sum(1))
            out.append(ds)
        return np.concatenate(out, axis=1)

    def rhs_vjp(self, t: float, y: np.ndarray, kbar: np.ndarray, with_score: bool):
        """Cotangent pull-back of :meth:`rhs`: returns (ybar (N, 4|7), theta_bar flat)."""
        y0 = self.y0
        n = y.shape[0]
        na = self.nuclei.shape[0]
        x = y[:, :3]
        dvec = x[:, None, :] - self.nuclei[None]
        r = np.sqrt((dvec * dvec).sum(-1))
        tp = y0 * t
        pre, wr = self._first_pre(tp, r)
        nj = 3 if with_score else 2
        f, f1, f2, cache = mlp_jet(self.params, pre.reshape(n * na, -1), wr, nj, keep=True)
        f = f.reshape(n, na); f1 = f1.reshape(n, na)
        ax = kbar[:, :3]; al = kbar[:, 3]
        alpha = (dvec * ax[:, None, :]).sum(-1)
        u = dvec / r[..., None]
        fb = y0 * (alpha - 3.0 * al[:, None])
        f1b = y0 * (-r * al[:, None])
        f2b = None
        dbar = y0 * f[..., None] * ax[:, None, :]
        rbar = y0 * (-f1 * al[:, None])
        sbar = None
        if with_score:
            f2 = f2.reshape(n, na)
            s = y[:, 4:7]; as_ = kbar[:, 4:7]
            q = (dvec * s[:, None, :]).sum(-1)
            beta = (as_ * s).sum(-1)
            gam = (u * as_[:, None, :]).sum(-1)
            amp = f1 * (q + 4.0) + f2 * r
            fb = fb - y0 * beta[:, None]
            f1b = f1b - y0 * (q + 4.0) * gam
            f2b = -y0 * r * gam
            rbar = rbar - y0 * f2 * gam
            qbar = -y0 * f1 * gam
            ubar = (-y0 * amp)[..., None] * as_[:, None, :]
            dbar = dbar + qbar[..., None] * s[:, None, :] + ubar / r[..., None]
            rbar = rbar - (ubar * u).sum(-1) / r
            sbar = (qbar[..., None] * dvec).sum(1) - y0 * f.sum(1)[:, None] * as_
        prebar, wrbar, (g_hidden, g_last) = mlp_jet_vjp(
            self.params, cache, fb.reshape(-1, 1), f1b.reshape(-1, 1),
            None if f2b is None else f2b.reshape(-1, 1))
        w1, b1 = self.params.hidden[0]
        prebar = prebar.reshape(n, na, -1)
        rbar = rbar + prebar @ w1[1]
        gw1 = np.zeros_like(w1)
        cbar = prebar.sum(0)                                  # (Na, H)
        gw1[0] = tp * cbar.sum(0)
        gw1[1] = wrbar + (prebar * r[..., None]).sum((0, 1))
        gw1[2:] = self.onehot.T @ cbar
        g_hidden[0] = (gw1, cbar.sum(0))
        xbar = (dbar + rbar[..., None] * u).sum(1)
        ybar = np.zeros_like(y)
        ybar[:, :3] = xbar
        if with_score:
            ybar[:, 4:7] = sbar
        gp = MLPParams(g_hidden, g_last)
        return ybar, flatten_params(gp)


@dataclass
class MLPField:
    """Gen_CNFSimpleMLP (cn_flows.py:166-182) for in_out_dim d = 1: h = MLP([t', x]); state rows
    [x, logp, (score)];  dx = y0 h, dlogp = -y0 h', dscore = -y0 h' s - y0 h''  (SURVEY.md 3.4)."""
    params: MLPParams
    bool_neg: bool = False
    d: int = 1

    @property
    def y0(self) -> float:
        return -1.0 if self.bool_neg else 1.0

    def rhs(self, t: float, y: np.ndarray, with_score: bool) -> np.ndarray:
        y0 = self.y0
        w1, b1 = self.params.hidden[0]
        pre = y[:, 0:1] * w1[1] + (y0 * t * w1[0] + b1)
        h, h1, h2 = mlp_jet(self.params, pre, w1[1], 3 if with_score else 2)
        out = [y0 * h, -y0 * h1]
        if with_score:
            out.append(-y0 * h1 * y[:, 2:3] - y0 * h2)
        return np.concatenate(out, axis=1)

    def rhs_vjp(self, t: float, y: np.ndarray, kbar: np.ndarray, with_score: bool):
        y0 = self.y0
        tp = y0 * t
        w1, b1 = self.params.hidden[0]
        x = y[:, 0:1]
        pre = x * w1[1] + (tp * w1[0] + b1)
        h, h1, h2, cache = mlp_jet(self.params, pre, w1[1], 3 if with_score else 2, keep=True)
        hb = y0 * kbar[:, 0:1]
        h1b = -y0 * kbar[:, 1:2]
        h2b = None
        ybar = np.zeros_like(y)
        if with_score:
            s = y[:, 2:3]; as_ = kbar[:, 2:3]
            h1b = h1b - y0 * s * as_
            h2b = -y0 * as_
            ybar[:, 2:3] = -y0 * h1 * as_
        prebar, wxbar, (g_hidden, g_last) = mlp_jet_vjp(self.params, cache, hb, h1b, h2b)
        ybar[:, 0:1] = prebar @ w1[1][:, None]
        gw1 = np.zeros_like(w1)
        gw1[0] = tp * prebar.sum(0)
        gw1[1] = wxbar + (prebar * x).sum(0)
        g_hidden[0] = (gw1, prebar.sum(0))
        return ybar, flatten_params(MLPParams(g_hidden, g_last))


# ----------------------------------------------------------------------------------------------
# integrators
# ----------------------------------------------------------------------------------------------

# [third-party] Dormand-Prince 5(4) tableau as used by jax.experimental.ode
_DP_ALPHA = np.array([1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0, 0.0])
_DP_BETA = np.array([
    [1 / 5, 0, 0, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0, 0, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656, 0, 0],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]])
_DP_CSOL = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0])
_DP_CERR = np.array([35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
                     -2187 / 6784 + 12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0])
_DP_CMID = np.array([6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
                     -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
                     -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2])


def odeint_dopri5(func: Callable[[np.ndarray, float], np.ndarray], y0: np.ndarray, t0: float, t1: float,
                  rtol: float = 1e-7, atol: float = 1e-7, mxstep: int = 100000):
    """[third-party] adaptive Dopri5 with FSAL as in jax.experimental.ode.odeint (call sites
    jax_ode.py:39-45,100-106): the WHOLE batch is one ODE system, the error norm is the mean of
    squared scaled errors over every entry, the step is never clipped to t1 (the result is read
    from the 4th-order interpolant).  Returns (y(t1), stats)."""
    shape = y0.shape
    y = y0.astype(np.float64).ravel()
    fun = lambda yy, tt: func(yy.reshape(shape), tt).ravel()
    f = fun(y, t0)
    nfe = 1
    # initial step (Hairer-Norsett-Wanner II.4, order = 4)
    scale = atol + np.abs(y) * rtol
    d0 = np.linalg.norm(y / scale); d1 = np.linalg.norm(f / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = fun(y + h0 * f, t0 + h0); nfe += 1
    d2 = np.linalg.norm((f1 - f) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    dt = min(100.0 * h0, h1)
    t = t0; last_t = t0
    coeff = [y] * 5
    n_acc = n_rej = 0
    while t < t1 and (n_acc + n_rej) < mxstep:
        k = np.zeros((7, y.size)); k[0] = f
        for i in range(1, 7):
            yi = y + dt * (_DP_BETA[i - 1, :i] @ k[:i])
            k[i] = fun(yi, t + dt * _DP_ALPHA[i - 1]); nfe += 1
        y1 = y + dt * (_DP_CSOL @ k)
        err = dt * (_DP_CERR @ k)
        tol = atol + rtol * np.maximum(np.abs(y), np.abs(y1))
        ratio = float(np.mean((err / tol) ** 2))
        # interpolant (fit_4th_order_polynomial)
        ymid = y + dt * (_DP_CMID @ k)
        dy0 = k[0]; dy1 = k[6]
        ca = -2 * dt * dy0 + 2 * dt * dy1 - 8 * y - 8 * y1 + 16 * ymid
        cb = 5 * dt * dy0 - 3 * dt * dy1 + 18 * y + 14 * y1 - 32 * ymid
        cc = -4 * dt * dy0 + dt * dy1 - 11 * y - 5 * y1 + 16 * ymid
        new_coeff = [ca, cb, cc, dt * dy0, y]
        # optimal_step_size(safety .9, ifactor 10, dfactor .2, order 5)
        if ratio == 0.0:
            new_dt = dt * 10.0
        else:
            dfac = 1.0 if ratio < 1.0 else 0.2
            factor = max(1.0 / 10.0, min(math.sqrt(ratio) ** (1.0 / 5.0) / 0.9, 1.0 / dfac))
            new_dt = dt / factor
        if ratio <= 1.0:
            last_t = t; t = t + dt; y = y1; f = k[6]; coeff = new_coeff; n_acc += 1
        else:
            n_rej += 1
        dt = new_dt
    rel = (t1 - last_t) / (t - last_t)
    out = coeff[0]
    for c in coeff[1:]:
        out = out * rel + c
    return out.reshape(shape), dict(accepted=n_acc, rejected=n_rej, nfe=nfe)


def odeint_rk4(func: Callable[[np.ndarray, float], np.ndarray], y0: np.ndarray, t0: float, t1: float,
               n_steps: int, keep: bool = False):
    """Classical RK4 with n_steps equal steps: the scheme the B200 kernels integrate with
    (BASELINE.json north_star: "fused fixed-step ODE integrator").  If ``keep`` also returns the
    stage-input states [(t_stage, y_stage)] * 4 per step that the discrete adjoint revisits."""
    h = (t1 - t0) / n_steps
    y = y0.astype(np.float64)
    stages = []
    for n in range(n_steps):
        t = t0 + n * h
        k1 = func(y, t)
        y2 = y + 0.5 * h * k1; k2 = func(y2, t + 0.5 * h)
        y3 = y + 0.5 * h * k2; k3 = func(y3, t + 0.5 * h)
        y4 = y + h * k3; k4 = func(y4, t + h)
        if keep:
            stages.append(((t, y), (t + 0.5 * h, y2), (t + 0.5 * h, y3), (t + h, y4)))
        y = y + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    return (y, stages) if keep else y


def rk4_adjoint(field, stages, t0: float, t1: float, ybar1: np.ndarray, with_score: bool):
    """Discrete adjoint of :func:`odeint_rk4` (discretise-then-optimise): exact gradient of the
    RK4 map.  Returns (ybar0, theta_bar)."""
    n_steps = len(stages)
    h = (t1 - t0) / n_steps
    ybar = ybar1.astype(np.float64).copy()
    tbar = None
    for n in range(n_steps - 1, -1, -1):
        (ta, ya), (tb, yb), (tc, yc), (td, yd) = stages[n]
        k4b = (h / 6.0) * ybar
        y4b, g4 = field.rhs_vjp(td, yd, k4b, with_score)
        k3b = (h / 3.0) * ybar + h * y4b
        y3b, g3 = field.rhs_vjp(tc, yc, k3b, with_score)
        k2b = (h / 3.0) * ybar + 0.5 * h * y3b
        y2b, g2 = field.rhs_vjp(tb, yb, k2b, with_score)
        k1b = (h / 6.0) * ybar + 0.5 * h * y2b
        y1b, g1 = field.rhs_vjp(ta, ya, k1b, with_score)
        ybar = ybar + y1b + y2b + y3b + y4b
        g = g1 + g2 + g3 + g4
        tbar = g if tbar is None else tbar + g
    return ybar, tbar


def neural_ode(field, batch: np.ndarray, t0: float, t1: float, method="dopri5", n_steps: int = 16):
    """jax_ode.py:9-49: returns (z_t1 (B,d), logp_t1 (B,1))."""
    d = field.d
    fn = lambda y, t: field.rhs(t, y, False)
    y1 = odeint_dopri5(fn, batch, t0, t1)[0] if method == "dopri5" else odeint_rk4(fn, batch, t0, t1, n_steps)
    return y1[:, :d], y1[:, d:d + 1]


def neural_ode_score(field, batch: np.ndarray, t0: float, t1: float, method="dopri5", n_steps: int = 16):
    """jax_ode.py:51-110: returns (z_t1 (B,d), logp_t1 (B,1), score_t1 (B,d))."""
    d = field.d
    fn = lambda y, t: field.rhs(t, y, True)
    y1 = odeint_dopri5(fn, batch, t0, t1)[0] if method == "dopri5" else odeint_rk4(fn, batch, t0, t1, n_steps)
    return y1[:, :d], y1[:, d:d + 1], y1[:, d + 1:]


# ----------------------------------------------------------------------------------------------
# functionals (functionals.py) -- per-sample integrands of shape (bs, 1), quirks reproduced as coded
# ----------------------------------------------------------------------------------------------

C_TF = (3.0 / 10.0) * (3.0 * math.pi ** 2) ** (2.0 / 3.0)       # functionals.py:101
C_TF1D = (math.pi * math.pi) / 24.0                               # functionals.py:131
LAMBDA_W = 0.2                                                    # functionals.py:70
HGH_H = dict(Zion=1.0, rloc=0.2, C1=-4.180237, C2=0.725075, C3=0.0, C4=0.0)       # hgh_pseudopotentials.py:8-9
HGH_O = dict(Zion=6.0, rloc=0.247621, C1=-16.580318, C2=2.395701, C3=0.0, C4=0.0)  # hgh_pseudopotentials.py:10-11


def thomas_fermi(den, score, Ne):
    """functionals.py:101-127: c Ne^{5/3} den^{2/3}."""
    return C_TF * (Ne ** (5.0 / 3.0)) * den ** (2.0 / 3.0)


def thomas_fermi_1D(den, score, Ne):
    """functionals.py:131-158: (pi^2/24) Ne^3 den^2."""
    return C_TF1D * (Ne ** 3) * den * den


def weizsacker(den, score, Ne):
    """functionals.py:70-97: (lambda0 Ne / 8) |score|^2 (ignores den)."""
    return (LAMBDA_W * Ne / 8.0) * (score * score).sum(-1, keepdims=True)


def lda(den, score, Ne):
    """functionals.py:391-417.  QUIRK (line 415): ``(3/pi)**1/3`` parses as ((3/pi)**1)/3 = 1/pi,
    not the cube root; reproduced as coded."""
    l = -(3.0 / 4.0) * (Ne ** (4.0 / 3.0)) * (3.0 / math.pi) ** 1 / 3
    return l * den ** (1.0 / 3.0)


def hartree_potential(x, xp, Ne, eps=1e-5):
    """functionals.py:463-486.  QUIRK: eps is added per component inside the sum (3*eps in 3-D)."""
    z = ((x - xp) * (x - xp) + eps).sum(-1, keepdims=True)
    return 0.5 * (Ne ** 2) / np.sqrt(z)


def soft_coulomb(x, xp, Ne):
    """functionals.py:439-461.  QUIRK: no 1/2 factor (unlike the 3-D Hartree)."""
    return (Ne ** 2) / np.sqrt(1.0 + (x - xp) * (x - xp))


def hartree_potential_mt(x, xp, Ne, alpha=0.5):
    """functionals.py:489-496: 0.5 Ne^2 (erf(a r)/r + erfc(a r)/r) = 0.5 Ne^2 / r (no eps).  Evaluated as coded, so an
    exactly coincident pair gives 0/0 + 1/0 = NaN like the reference (the CUDA kernel returns +inf there)."""
    from scipy.special import erf, erfc
    r = np.sqrt(((x - xp) ** 2).sum(-1, keepdims=True))
    with np.errstate(divide="ignore", invalid="ignore"):
        return 0.5 * (Ne ** 2) * (erf(alpha * r) / r + erfc(alpha * r) / r)


def nuclei_potential(x, Ne, coords, z, eps=1e-4):
    """functionals.py:556-587.  QUIRK: eps is added to r, not r^2."""
    r = np.sqrt(((x[:, None, :] - coords[None]) ** 2).sum(-1)) + eps
    return -Ne * (z[None] / r).sum(-1, keepdims=True)


def attraction(x, R, Z_alpha, Z_beta, Ne):
    """functionals.py:523-549 (1-D; Z_alpha sits at -R/2)."""
    return Ne * (-Z_alpha / np.sqrt(1.0 + (x + R / 2.0) ** 2) - Z_beta / np.sqrt(1.0 + (x - R / 2.0) ** 2))


def nuclei_potential_hgh(x, Ne, coords, z, per_species: bool = False):
    """functionals.py:617-655.  QUIRK (lines 619-625,649): the hydrogen parameters are used for
    every atom.  ``per_species=True`` (NOT the reference behaviour) picks H/O tables by Z."""
    from scipy.special import erf
    r = np.sqrt(((x[:, None, :] - coords[None]) ** 2).sum(-1))
    tot = np.zeros((x.shape[0], 1))
    for i in range(coords.shape[0]):
        pp = HGH_O if (per_species and int(round(z[i])) == 8) else HGH_H
        rr = r[:, i:i + 1] / pp["rloc"]
        rr2 = rr * rr
        v0 = (-pp["Zion"] / r[:, i:i + 1]) * erf(rr / math.sqrt(2.0))
        v1 = np.exp(-0.5 * rr2)
        v2 = pp["C1"] + pp["C2"] * rr2 + pp["C3"] * rr ** 4 + pp["C4"] * rr ** 6
        tot += v0 + v1 * v2
    return Ne * tot


def b88_x_e(den, score, Ne):
    """functionals.py:335-389 (SURVEY 8f "next")."""
    clip = 1e-30; beta = 0.0042
    den = np.clip(den, clip, None)
    log_den = np.log2(den)
    g2 = (score * score).sum(-1, keepdims=True) * den * den
    log_x = np.log2(np.clip(g2, clip, None)) / 2 - 4.0 / 3.0 * log_den
    xs = 2.0 ** log_x
    e = -(beta * 2.0 ** (4 * log_den / 3 + 2 * log_x - np.log2(1 + 6 * beta * xs * np.arcsinh(xs))))
    return e * Ne ** (2.0 / 3.0)


def correlation_vwn_c_e(den, Ne):
    """functionals.py:185-245 (SURVEY 8f "next")."""
    A = 0.0621814; b = 3.72744; c = 12.9352; x0 = -0.10498; clip = 1e-30
    den = np.where(den > clip, den, 0.0)
    log_den = np.log2(np.clip(den, clip, None))
    log_rs = math.log2((3 / (4 * math.pi)) ** (1 / 3)) - log_den / 3.0
    log_x = log_rs / 2
    x = 2.0 ** log_x
    X = 2.0 ** (2.0 * log_x) + 2.0 ** (log_x + math.log2(b)) + c
    X0 = x0 ** 2 + b * x0 + c
    Q = math.sqrt(4 * c - b ** 2)
    e = A / 2.0 * (2.0 * np.log(x) - np.log(X) + 2.0 * b / Q * np.arctan(Q / (2.0 * x + b))
                   - b * x0 / X0 * (np.log((x - x0) ** 2.0 / X) + 2.0 * (2.0 * x0 + b) / Q * np.arctan(Q / (2.0 * x + b))))
    return Ne * e


def correlation_pw92_c_e(den, Ne):
    """functionals.py:247-292 (SURVEY 8f "next")."""
    clip = 1e-30
    A_ = 0.031091; alpha1 = 0.21370; beta1 = 7.5957; beta2 = 3.5876; beta3 = 1.6382; beta4 = 0.49294
    log_den = np.log2(np.clip(den.sum(axis=1, keepdims=True), clip, None))
    log_rs = math.log2((3 / (4 * math.pi)) ** (1 / 3)) - log_den / 3.0
    brs_1_2 = 2 ** (log_rs / 2 + math.log2(beta1))
    ars = 2 ** (log_rs + math.log2(alpha1))
    brs = 2 ** (log_rs + math.log2(beta2))
    brs_3_2 = 2 ** (3 * log_rs / 2 + math.log2(beta3))
    brs2 = 2 ** (2 * log_rs + math.log2(beta4))
    e = -2 * A_ * (1 + ars) * np.log(1 + (1 / (2 * A_)) / (brs_1_2 + brs + brs_3_2 + brs2))
    return Ne * e


def exchange_correlation_one_dimensional(den, Ne):
    """functionals.py:294-332 (SURVEY 8f "next")."""
    rs = 1 / (2 * Ne * den)
    a0 = -0.8862269; b0 = -2.1414101; c0 = 0.4721355; d0 = 2.81423; e0 = 0.529891; f0 = 0.458513
    g0 = -0.202642; h0 = 0.470876; alpha0 = 0.104435; beta0 = 4.11613
    f1 = (a0 + b0 * rs + c0 * rs ** 2) / (1 + d0 * rs + e0 * rs ** 2 + f0 * rs ** 3)
    f2 = g0 * rs * np.log(rs + alpha0 * rs ** beta0) / (1 + h0 * rs ** 2)
    return Ne * (f1 + f2)


def _kinetic(name: str):
    """functionals.py:21-43 ("" = term switched off: zeros)."""
    n = name.lower()
    if not n:
        return lambda den, score, Ne: np.zeros_like(den)
    if n in ("tf", "thomas_fermi"):
        return thomas_fermi
    if n in ("tf1d", "thomas_fermi_1d"):
        return thomas_fermi_1D
    if n in ("w", "weizsacker", "w1d", "weizsacker1d"):
        return weizsacker
    if n in ("tf-w", "thomas_fermi_weizsacker"):
        return lambda *a: thomas_fermi(*a) + weizsacker(*a)
    if n in ("tfw_1d", "thomas_fermi_weizsacker_1d"):
        return lambda *a: thomas_fermi_1D(*a) + weizsacker(*a)
    raise KeyError(name)


def _hartree(name: str):
    """functionals.py:425-436."""
    n = name.lower()
    if not n:
        return lambda x, xp, Ne: np.zeros((x.shape[0], 1))
    if n in ("hartree_potential", "hartree"):
        return hartree_potential
    if n in ("soft_coulomb", "softc"):
        return soft_coulomb
    if n in ("hartree_mt",):
        return hartree_potential_mt
    raise KeyError(name)


def _nuclear(name: str):
    """functionals.py:504-521; 3-D variants take (x, Ne, mol_info={'coords','z'})."""
    n = name.lower()
    if n in ("nuclei_potential", "nuclei"):
        return lambda x, Ne, mol: nuclei_potential(x, Ne, mol["coords"], mol["z"])
    if n in ("attraction", "attr"):
        return attraction
    if n in ("hgh", "nuclei_potential_hgh"):
        return lambda x, Ne, mol: nuclei_potential_hgh(x, Ne, mol["coords"], mol["z"])
    raise KeyError(name)


def _exchange_correlation(name: str):
    """functionals.py:164-183."""
    n = name.lower()
    if n in ("lda", "local_density_approximation"):
        return lda
    if n in ("xc_1d", "exchange_correlation_1d"):
        return exchange_correlation_one_dimensional
    if n in ("vwn_c_e", "correlation_vwn_c_e"):
        return correlation_vwn_c_e
    if n in ("pw92_c_e", "correlation_pw92_c_e"):
        return correlation_pw92_c_e
    if n == "b88_x_e":
        return b88_x_e
    if n == "dirac_b88_x_e":
        return lambda *a: lda(*a) + b88_x_e(*a)
    raise KeyError(name)


# ----------------------------------------------------------------------------------------------
# loss glue (H20_mol_ofdft_min_equiv_flows.py:150-173, LiH.py:118-138) and its gradient
# ----------------------------------------------------------------------------------------------

@dataclass
class EnergySpec:
    """Functional selection of one driver run (argparse names, H20_...py:225-249 / LiH.py:211-223)."""
    Ne: float
    kin: str = "tf-w"
    hart: str = "hartree"
    nuc: str = "nuclei_potential"
    x: str = "lda"            # "" = off
    c: str = ""               # c-type functional (den, Ne); "" = off
    coords: np.ndarray = None  # (Na,3) for 3-D nuclear terms
    z: np.ndarray = None
    R: float = 0.0            # 1-D attraction
    Z_alpha: float = 0.0
    Z_beta: float = 0.0


def hartree_allpairs(x, xp, Ne, eps=1e-5):
    """North-star generalisation of the paired estimator (functionals.py:463-486): the same kernel averaged over ALL
    cross-half pairs.  Returns (mean energy, d/dx, d/dxp)."""
    diff = x[:, None, :] - xp[None]
    zz = (diff * diff + eps).sum(-1)
    n, m = zz.shape
    e = 0.5 * Ne ** 2 * (1.0 / np.sqrt(zz)).mean()
    gk = -0.5 * Ne ** 2 * diff / zz[..., None] ** 1.5 / (n * m)
    return e, gk.sum(1), -gk.sum(0)


def softc_allpairs(x, xp, Ne):
    """The 1-D soft-Coulomb kernel (functionals.py:439-461, no 1/2 factor) averaged over ALL cross-half pairs.
    Returns (mean energy, d/dx, d/dxp)."""
    diff = x[:, None, :] - xp[None]
    zz = 1.0 + (diff * diff).sum(-1)
    n, m = zz.shape
    e = Ne ** 2 * (1.0 / np.sqrt(zz)).mean()
    gk = -(Ne ** 2) * diff / zz[..., None] ** 1.5 / (n * m)
    return e, gk.sum(1), -gk.sum(0)


def _allpairs_fn(name: str):
    return {"hartree_allpairs": hartree_allpairs, "softc_allpairs": softc_allpairs}.get(name.lower())


def energy_terms(spec: EnergySpec, y1: np.ndarray, d: int):
    """Per-term means [kin, vnuc, hart, x, c] and total, from the t1 state (2*bs rows).
    Only the first half's den/score/x enter the local functionals; the second half supplies xp
    (H20_...py:152-156)."""
    bs = y1.shape[0] // 2
    x, xp = y1[:bs, :d], y1[bs:, :d]
    den = np.exp(y1[:bs, d:d + 1])
    score = y1[:bs, d + 1:2 * d + 1]
    e_t = _kinetic(spec.kin)(den, score, spec.Ne)
    allpairs = _allpairs_fn(spec.hart)
    if not spec.hart:
        e_h = np.zeros_like(den)
    else:
        e_h = np.full_like(den, allpairs(x, xp, spec.Ne)[0]) if allpairs else _hartree(spec.hart)(x, xp, spec.Ne)
    if spec.nuc.lower() in ("attraction", "attr"):
        e_n = attraction(x, spec.R, spec.Z_alpha, spec.Z_beta, spec.Ne)
    else:
        e_n = _nuclear(spec.nuc)(x, spec.Ne, {"coords": spec.coords, "z": spec.z})
    e_x = _exchange_correlation(spec.x)(den, score, spec.Ne) if spec.x else np.zeros_like(e_t)
    e_c = _exchange_correlation(spec.c)(den, spec.Ne) if spec.c else np.zeros_like(e_t)
    terms = np.array([e_t.mean(), e_n.mean(), e_h.mean(), e_x.mean(), e_c.mean()])
    return float((e_t + e_n + e_h + e_x + e_c).mean()), terms


def energy_terms_grad(spec: EnergySpec, y1: np.ndarray, d: int, fd: float = 1e-6):
    """Cotangent d(energy)/d(y1) by central differences row-wise (every functional is row-local or
    row-paired, so one perturbation per column suffices).  Oracle only -- O(columns) evaluations."""
    bs = y1.shape[0] // 2
    g = np.zeros_like(y1)

    def per_row(yy):
        x, xp = yy[:bs, :d], yy[bs:, :d]
        den = np.exp(yy[:bs, d:d + 1]); score = yy[:bs, d + 1:2 * d + 1]
        e = _kinetic(spec.kin)(den, score, spec.Ne) + _hartree(spec.hart)(x, xp, spec.Ne)
        if spec.nuc.lower() in ("attraction", "attr"):
            e = e + attraction(x, spec.R, spec.Z_alpha, spec.Z_beta, spec.Ne)
        else:
            e = e + _nuclear(spec.nuc)(x, spec.Ne, {"coords": spec.coords, "z": spec.z})
        if spec.x:
            e = e + _exchange_correlation(spec.x)(den, score, spec.Ne)
        if spec.c:
            e = e + _exchange_correlation(spec.c)(den, spec.Ne)
        return e[:, 0]

    for half in (0, 1):
        for col in range(y1.shape[1]):
            yp = y1.copy(); ym = y1.copy()
            rows = slice(0, bs) if half == 0 else slice(bs, 2 * bs)
            hstep = fd * np.maximum(1.0, np.abs(y1[rows, col]))
            yp[rows, col] += hstep; ym[rows, col] -= hstep
            g[rows, col] = (per_row(yp) - per_row(ym)) / (2 * hstep) / bs
    return g


def energy_terms_vjp(spec: EnergySpec, y1: np.ndarray, d: int):
    """Analytic cotangent d(mean energy)/d(y1) for the north-star functionals (TF/TF1D, W, lda,
    Hartree / soft-Coulomb / MT, nuclear / HGH / attraction).  XC "next" terms are not covered."""
    from scipy.special import erf
    if spec.c or (spec.x and spec.x.lower() not in ("lda", "local_density_approximation")):
        return energy_terms_grad(spec, y1, d)      # "next" XC terms: row-wise central differences
    bs = y1.shape[0] // 2
    Ne = spec.Ne
    g = np.zeros_like(y1)
    x, xp = y1[:bs, :d], y1[bs:, :d]
    logp = y1[:bs, d:d + 1]
    den = np.exp(logp); score = y1[:bs, d + 1:2 * d + 1]
    kin = spec.kin.lower()
    if kin in ("tf", "thomas_fermi", "tf-w", "thomas_fermi_weizsacker"):
        g[:bs, d:d + 1] += (2.0 / 3.0) * thomas_fermi(den, score, Ne)
    if kin in ("tf1d", "thomas_fermi_1d", "tfw_1d", "thomas_fermi_weizsacker_1d"):
        g[:bs, d:d + 1] += 2.0 * thomas_fermi_1D(den, score, Ne)
    if kin in ("w", "weizsacker", "w1d", "weizsacker1d", "tf-w", "thomas_fermi_weizsacker", "tfw_1d",
               "thomas_fermi_weizsacker_1d"):
        g[:bs, d + 1:2 * d + 1] += (LAMBDA_W * Ne / 4.0) * score
    if spec.x:
        assert spec.x.lower() in ("lda", "local_density_approximation")
        g[:bs, d:d + 1] += (1.0 / 3.0) * lda(den, score, Ne)
    assert not spec.c
    hn = spec.hart.lower()
    diff = x - xp
    if not hn:
        gx = np.zeros_like(diff)
    elif _allpairs_fn(hn) is not None:
        _, gxa, gxpa = _allpairs_fn(hn)(x, xp, Ne)
        g[:bs, :d] += gxa * bs          # the common 1/bs below is undone: the pair mean is already normalised
        g[bs:, :d] += gxpa * bs
        gx = np.zeros_like(diff)
    elif hn in ("hartree_potential", "hartree"):
        zz = (diff * diff + 1e-5).sum(-1, keepdims=True)
        gx = -0.5 * Ne ** 2 * diff / zz ** 1.5
    elif hn in ("soft_coulomb", "softc"):
        gx = -(Ne ** 2) * diff / (1.0 + diff * diff) ** 1.5
    else:
        zz = (diff * diff).sum(-1, keepdims=True)
        gx = -0.5 * Ne ** 2 * diff / zz ** 1.5
    g[:bs, :d] += gx
    g[bs:, :d] -= gx
    nn = spec.nuc.lower()
    if nn in ("attraction", "attr"):
        a = x + spec.R / 2.0; b = x - spec.R / 2.0
        g[:bs, :d] += Ne * (spec.Z_alpha * a / (1 + a * a) ** 1.5 + spec.Z_beta * b / (1 + b * b) ** 1.5)
    else:
        dv = x[:, None, :] - spec.coords[None]
        r = np.sqrt((dv * dv).sum(-1))
        if nn in ("nuclei_potential", "nuclei"):
            dvdr = Ne * spec.z[None] / (r + 1e-4) ** 2
        else:
            pp = HGH_H
            rl = pp["rloc"]; rr = r / rl; rr2 = rr * rr
            ex = np.exp(-0.5 * rr2)
            poly = pp["C1"] + pp["C2"] * rr2 + pp["C3"] * rr2 ** 2 + pp["C4"] * rr2 ** 3
            dpoly = (2 * pp["C2"] * rr + 4 * pp["C3"] * rr ** 3 + 6 * pp["C4"] * rr ** 5) / rl
            dv0 = pp["Zion"] / r ** 2 * erf(rr / math.sqrt(2)) - (pp["Zion"] / r) * math.sqrt(2 / math.pi) / rl * ex
            dvdr = Ne * (dv0 + ex * (dpoly - rr / rl * poly))
        g[:bs, :d] += (dvdr[..., None] * dv / r[..., None]).sum(1)
    return g / bs


def loss_and_grad(field, spec: EnergySpec, batch: np.ndarray, t0=0.0, t1=1.0, n_steps: int = 16):
    """value_and_grad(loss) (H20_...py:175-178) with the fixed-step RK4 + discrete adjoint the
    kernels implement, in float64.  Returns (energy, terms[5], theta_bar flat, y1)."""
    d = field.d
    fn = lambda y, t: field.rhs(t, y, True)
    y1, stages = odeint_rk4(fn, batch, t0, t1, n_steps, keep=True)
    energy, terms = energy_terms(spec, y1, d)
    ybar1 = energy_terms_vjp(spec, y1, d)
    _, tbar = rk4_adjoint(field, stages, t0, t1, ybar1, True)
    return energy, terms, tbar, y1


def loss_adaptive(field, spec: EnergySpec, batch: np.ndarray, t0=0.0, t1=1.0):
    """loss (H20_...py:150-173) with the reference's adaptive Dopri5 (rtol=atol=1e-7)."""
    fn = lambda y, t: field.rhs(t, y, True)
    y1, stats = odeint_dopri5(fn, batch, t0, t1)
    energy, terms = energy_terms(spec, y1, field.d)
    return energy, terms, y1, stats


# ----------------------------------------------------------------------------------------------
# priors / batch generators (input producers; utils.py:76-143, promolecular_distrax.py:40-81)
# ----------------------------------------------------------------------------------------------

def promolecular_logp_score(x: np.ndarray, coords: np.ndarray, z: np.ndarray):
    """ProMolecularDensity.log_prob / score: mixture of unit-variance Gaussians at the nuclei with
    weights Z_i / sum Z (promolecular_distrax.py:52-62,86-89)."""
    w = z / np.abs(z).sum()
    dv = x[:, None, :] - coords[None]
    lg = -0.5 * (dv * dv).sum(-1) - 1.5 * math.log(2 * math.pi)
    m = lg.max(1, keepdims=True)
    pe = np.exp(lg - m) * w[None]
    tot = pe.sum(1, keepdims=True)
    logp = np.log(tot) + m
    resp = pe / tot
    score = -(resp[..., None] * dv).sum(1)
    return logp, score


def batch_generator_3d(rng: np.random.Generator, bs: int, coords: np.ndarray, z: np.ndarray) -> np.ndarray:
    """utils.py:76-108: two independent draws of bs samples, rows [x(3), logp0, score0(3)],
    concatenated on axis 0 -> (2*bs, 7)."""
    out = []
    w = z / np.abs(z).sum()
    for _ in range(2):
        comp = rng.choice(len(z), size=bs, p=w)
        x = coords[comp] + rng.standard_normal((bs, 3))
        logp, score = promolecular_logp_score(x, coords, z)
        out.append(np.concatenate([x, logp, score], 1))
    return np.concatenate(out, 0)


def batch_generator_1d(rng: np.random.Generator, bs: int) -> np.ndarray:
    """utils.py:110-143 with prior MultivariateNormalDiag(0,1) (LiH.py:78): rows [x, logp0, score0]."""
    out = []
    for _ in range(2):
        x = rng.standard_normal((bs, 1))
        out.append(np.concatenate([x, -0.5 * x * x - 0.5 * math.log(2 * math.pi), -x], 1))
    return np.concatenate(out, 0)


# ----------------------------------------------------------------------------------------------
# molecules (utils.py:276-326 and the drivers) + the one-hot rules
# ----------------------------------------------------------------------------------------------

def jax_one_hot_float(z: np.ndarray) -> np.ndarray:
    """``jax.nn.one_hot(z_float, unique(z)+1)`` (H20_...py:83-84): class index = the atomic number
    itself, so any Z >= n_types maps to an all-zero row (QUIRK: O, Li -> zeros)."""
    n = np.unique(z).shape[0] + 1
    oh = np.zeros((len(z), n))
    for i, zi in enumerate(z):
        if 0 <= int(zi) < n:
            oh[i, int(zi)] = 1.0
    return oh


def index_one_hot(z: np.ndarray) -> np.ndarray:
    """utils.py:251-274 ``one_hot_encode`` (used by OFDFT_NF.py:73): index into sorted unique Z."""
    zu = np.unique(z)
    oh = np.zeros((len(z), len(zu)))
    for i, zi in enumerate(z):
        oh[i, int(np.where(zu == zi)[0][0])] = 1.0
    return oh


def molecule(name: str):
    """Geometry (bohr), Z, Ne, one-hot as the corresponding driver builds them."""
    n = name.upper()
    if n == "H2":      # H2_mol_ofdft_min_equiv_flows.py:108-118
        c = np.array([[0., 0., -1.4008538753 / 2], [0., 0., 1.4008538753 / 2]]) * BOHR
        z = np.array([1., 1.]); return dict(coords=c, z=z, Ne=2, onehot=jax_one_hot_float(z))
    if n == "LIH":     # LiH_mol_ofdft_min_equiv_flows.py:115-126
        c = np.array([[0., 0., -1.5949 / 2], [0., 0., 1.5949 / 2]]) * BOHR
        z = np.array([3., 1.]); return dict(coords=c, z=z, Ne=4, onehot=jax_one_hot_float(z))
    if n == "H2O":     # H20_mol_ofdft_min_equiv_flows.py:70-84
        c = np.array([[0.0, 0.0, 0.1189120], [0.0, 0.7612710, -0.4756480], [0.0, -0.7612710, -0.4756480]]) * BOHR
        z = np.array([8., 1., 1.]); return dict(coords=c, z=z, Ne=10, onehot=jax_one_hot_float(z))
    if n == "C6H6":    # utils.py:310-326 + OFDFT_NF.py:73
        c = np.array([[-0.6984022192, 1.2096794375, 0.0001298085], [0.6983971652, 1.2096794375, 0.0001298085],
                      [1.3968148123, 0.0000100819, 0.0001298085], [0.6983970907, -1.2096631450, -0.0001577731],
                      [-0.6984347101, -1.2097236297, -0.0001117213], [-1.3967427862, -0.0000316846, -0.0000531302],
                      [2.4989957176, 0.0000352657, 0.0001433155], [1.2492115488, 2.1643070022, 0.0000608098],
                      [-1.2497352350, 2.1640355443, 0.0000873078], [1.2495192781, -2.1641211865, -0.0004652006],
                      [-1.2494488074, -2.1642720181, -0.0003281198], [-2.4988922798, 0.0006052813, -0.0002941369]]) * BOHR
        z = np.array([6.] * 6 + [1.] * 6); return dict(coords=c, z=z, Ne=42, onehot=index_one_hot(z))
    raise KeyError(name)


# ----------------------------------------------------------------------------------------------
# optimiser step (H20_mol_ofdft_min_equiv_flows.py:111-116) -- [third-party] optax semantics, restated from memory
# ----------------------------------------------------------------------------------------------

def rmsprop_clip_step(theta, grad, nu, lr, decay=0.9, eps=1e-8, max_norm=1.0):
    """optax.chain(clip_by_global_norm(max_norm), rmsprop(lr)): returns (theta', nu')."""
    gn = np.sqrt((grad * grad).sum())
    g = grad * (max_norm / gn if gn > max_norm else 1.0)
    nu2 = decay * nu + (1.0 - decay) * g * g
    return theta - lr * g / np.sqrt(nu2 + eps), nu2


def adam_clip_step(theta, grad, mu, nu, step: int, lr, b1=0.9, b2=0.999, eps=1e-8, max_norm=1.0):
    """[third-party] optax.chain(clip_by_global_norm(max_norm), adam(lr)) as OFDFT_NF.py:104-108 builds it, restated from the
    published algorithm (optax is not in the image): bias-corrected moments, eps OUTSIDE the square root (eps_root = 0);
    ``step`` is the 1-based update count.  Returns (theta', mu', nu')."""
    gn = np.sqrt((grad * grad).sum())
    g = grad * (max_norm / gn if gn > max_norm else 1.0)
    mu2 = b1 * mu + (1.0 - b1) * g
    nu2 = b2 * nu + (1.0 - b2) * g * g
    mhat = mu2 / (1.0 - b1 ** step)
    vhat = nu2 / (1.0 - b2 ** step)
    return theta - lr * mhat / (np.sqrt(vhat) + eps), mu2, nu2


def ema_update(values, state, step: int, decay: float = 0.99):
    """[third-party] optax.ema(decay, debias=True), restated from memory (call sites H20_mol_ofdft_min_equiv_flows.py:117-118,
    193-196): returns (debiased average, new biased state); ``step`` is the 1-based update count."""
    new = decay * np.asarray(state, dtype=np.float64) + (1.0 - decay) * np.asarray(values, dtype=np.float64)
    return new / (1.0 - decay ** step), new
