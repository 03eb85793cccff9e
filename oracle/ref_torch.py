"""Autodiff transliteration of the reference hot path (torch.func, CPU, float64).

TEST INFRASTRUCTURE ONLY (see oracle/ofdft_oracle.py header; PARITY UNPINNED: JAX is not
installable here, so the reference itself cannot run).  This file deliberately does NOT use the
closed forms of ofdft_oracle.py: it follows the reference source construct by construct --
``vmap(jacrev(f))`` for the trace, ``value_and_grad`` + ``vjp`` for the score dynamics, reverse
mode through the solver for the parameter gradient -- so that the two restatements pin each
other.  It is also the "restated reference" CPU baseline of bench.py (what XLA-CPU would execute:
generic nested AD, adaptive Dopri5 at rtol=atol=1e-7, float64).

torch is used here as a CPU autodiff engine only; it is not part of the product path.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
from torch.func import grad, jacrev, vjp, vmap

DT = torch.float64


def params_from_flat(theta, n_in: int, features: Sequence[int], n_out: int = 1) -> Dict[str, torch.Tensor]:
    """Flat theta (ravel_pytree key order) -> {'last_layer.bias', 'last_layer.kernel',
    'layers_i.bias', 'layers_i.kernel'}."""
    theta = torch.as_tensor(theta, dtype=DT)
    o = 0
    p = {}
    hl = features[-1]
    p["last_layer.bias"] = theta[o:o + n_out]; o += n_out
    p["last_layer.kernel"] = theta[o:o + hl * n_out].reshape(hl, n_out); o += hl * n_out
    prev = n_in
    for i, h in enumerate(features):
        p[f"layers_{i}.bias"] = theta[o:o + h]; o += h
        p[f"layers_{i}.kernel"] = theta[o:o + prev * h].reshape(prev, h); o += prev * h
        prev = h
    assert o == theta.numel()
    return p


def flat_from_params(p: Dict[str, torch.Tensor], n_layers: int) -> torch.Tensor:
    parts = [p["last_layer.bias"].reshape(-1), p["last_layer.kernel"].reshape(-1)]
    for i in range(n_layers):
        parts += [p[f"layers_{i}.bias"].reshape(-1), p[f"layers_{i}.kernel"].reshape(-1)]
    return torch.cat(parts)


def _dense_stack(p, z, n_layers):
    # NN.__call__ / SimpleMLP.__call__: Dense -> tanh per hidden layer, then last_layer
    for i in range(n_layers):
        z = torch.tanh(z @ p[f"layers_{i}.kernel"] + p[f"layers_{i}.bias"])
    return z @ p["last_layer.kernel"] + p["last_layer.bias"]


class GenEqvFlow:
    """equiv_flows.py:36-127 (NN, RadialMLP, EqvFlow, Gen_EqvFlow)."""

    def __init__(self, features, xyz_nuclei, z_one_hot, bool_neg=False):
        self.n_layers = len(features)
        self.nuclei = torch.as_tensor(xyz_nuclei, dtype=DT)[:, None]        # equiv_flows.py:66
        self.z_ = torch.as_tensor(z_one_hot, dtype=DT)
        self.y0 = -1.0 if bool_neg else 1.0                                 # equiv_flows.py:119-122
        self.d = 3

    def _nn(self, p, t, xi_norm, zi_one_hot):
        z = torch.hstack((t.reshape(1), xi_norm, zi_one_hot))               # equiv_flows.py:51
        return _dense_stack(p, z, self.n_layers)

    def _radial(self, p, t, samples):
        z = samples[None] - self.nuclei                                     # (Na,1,3)  equiv_flows.py:76
        z_norm = torch.linalg.norm(z, dim=-1)                               # (Na,1)
        x = vmap(lambda rn, oh: self._nn(p, t, rn, oh))(z_norm, self.z_)    # nn.vmap over nuclei, shared params
        return torch.einsum('ijk,ij->k', z, x)                              # equiv_flows.py:79

    def _cnf(self, p, t, states):
        z = states[:, :3]
        f = lambda zz: self._radial(p, t, zz)
        df_dz = vmap(jacrev(f))(z)                                          # equiv_flows.py:101
        dlogp = -1.0 * torch.einsum('bii->b', df_dz)
        dz = vmap(f)(z)
        return torch.cat((dz, dlogp[:, None]), 1)

    def apply(self, p, t, states):
        t = torch.as_tensor(t, dtype=DT)
        return self.y0 * self._cnf(p, self.y0 * t, states)                  # equiv_flows.py:125-127


class GenCNFSimpleMLP:
    """cn_flows.py:117-182 (SimpleMLP, CNFSimpleMLP, Gen_CNFSimpleMLP)."""

    def __init__(self, in_out_dim, features, bool_neg=False):
        self.n_layers = len(features)
        self.d = in_out_dim
        self.y0 = -1.0 if bool_neg else 1.0

    def _net(self, p, t, samples):
        z = torch.cat((t.reshape(1), samples))                              # cn_flows.py:133-134 ([t, x])
        return _dense_stack(p, z, self.n_layers)

    def _cnf(self, p, t, states):
        d = self.d
        z = states[:, :d]
        f = lambda zz: self._net(p, t, zz)
        df_dz = vmap(jacrev(f))(z)
        dlogp = -1.0 * torch.einsum('bii->b', df_dz)
        dz = vmap(f)(z)
        return torch.cat((dz, dlogp[:, None]), 1)

    def apply(self, p, t, states):
        t = torch.as_tensor(t, dtype=DT)
        return self.y0 * self._cnf(p, self.y0 * t, states)


def evol_score(model, p, t, states):
    """jax_ode.py:80-98 ``_evol_fn_i`` vmapped over rows."""
    d = model.d

    def one(state):
        st = state[None]
        xs, score = st[:, :-d], st[:, -d:]
        x_in = xs[:, :-1]

        def _f_div(x):
            full = torch.cat((x, torch.zeros(1, 1, dtype=DT)), 1)
            return model.apply(p, t, full)[:, -1:].sum()

        def _f_dx(x):
            full = torch.cat((x, torch.zeros(1, 1, dtype=DT)), 1)
            return model.apply(p, t, full)[:, :-1]

        div = _f_div(x_in)
        grad_div = grad(_f_div)(x_in)
        dx, f_vjp = vjp(_f_dx, x_in)
        score_vjp = f_vjp(score)[0]
        dscore = -score_vjp + grad_div
        return torch.cat((dx, div.reshape(1, 1), dscore), 1).reshape(-1)

    return vmap(one)(states)


# [third-party] Dormand-Prince tableau (jax.experimental.ode) -- see ofdft_oracle.odeint_dopri5
_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_BETA = [[1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9],
         [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
         [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
         [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
_CSOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_CERR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
         -2187 / 6784 + 12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0]
_CMID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
         187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def odeint_dopri5(func, y0, t0, t1, rtol=1e-7, atol=1e-7, mxstep=100000):
    """Adaptive Dopri5 of jax.experimental.ode, differentiable by back-propagation through the
    accepted/rejected step sequence (step sizes are treated as constants).  The reference instead
    differentiates with the continuous adjoint (``_odeint_rev``); both converge to the exact
    gradient at O(tol)."""
    y = y0
    f = func(y, t0)
    nfe = 1
    with torch.no_grad():
        scale = atol + y.abs() * rtol
        d0 = torch.linalg.norm((y / scale).reshape(-1)).item()
        d1 = torch.linalg.norm((f / scale).reshape(-1)).item()
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        f1 = func(y + h0 * f, t0 + h0); nfe += 1
        d2 = torch.linalg.norm(((f1 - f) / scale).reshape(-1)).item() / h0
        h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
        dt = min(100.0 * h0, h1)
    t = t0; last_t = t0
    coeff = None
    n_acc = n_rej = 0
    while t < t1 and (n_acc + n_rej) < mxstep:
        k = [f]
        for i in range(6):
            yi = y + dt * sum(b * kk for b, kk in zip(_BETA[i], k) if b != 0)
            k.append(func(yi, t + dt * _ALPHA[i])); nfe += 1
        y1 = y + dt * sum(c * kk for c, kk in zip(_CSOL, k) if c != 0)
        with torch.no_grad():
            err = dt * sum(c * kk for c, kk in zip(_CERR, k) if c != 0)
            tol = atol + rtol * torch.maximum(y.abs(), y1.abs())
            ratio = float(((err / tol) ** 2).mean())
        if ratio == 0.0:
            new_dt = dt * 10.0
        else:
            dfac = 1.0 if ratio < 1.0 else 0.2
            new_dt = dt / max(0.1, min(math.sqrt(ratio) ** 0.2 / 0.9, 1.0 / dfac))
        if ratio <= 1.0:
            ymid = y + dt * sum(c * kk for c, kk in zip(_CMID, k) if c != 0)
            dy0, dy1 = k[0], k[6]
            ca = -2 * dt * dy0 + 2 * dt * dy1 - 8 * y - 8 * y1 + 16 * ymid
            cb = 5 * dt * dy0 - 3 * dt * dy1 + 18 * y + 14 * y1 - 32 * ymid
            cc = -4 * dt * dy0 + dt * dy1 - 11 * y - 5 * y1 + 16 * ymid
            coeff = [ca, cb, cc, dt * dy0, y]
            last_t = t; t = t + dt; y = y1; f = k[6]; n_acc += 1
        else:
            n_rej += 1
        dt = new_dt
    rel = (t1 - last_t) / (t - last_t)
    out = coeff[0]
    for c in coeff[1:]:
        out = out * rel + c
    return out, dict(accepted=n_acc, rejected=n_rej, nfe=nfe)


def odeint_rk4(func, y0, t0, t1, n_steps):
    h = (t1 - t0) / n_steps
    y = y0
    for n in range(n_steps):
        t = t0 + n * h
        k1 = func(y, t)
        k2 = func(y + 0.5 * h * k1, t + 0.5 * h)
        k3 = func(y + 0.5 * h * k2, t + 0.5 * h)
        k4 = func(y + h * k3, t + h)
        y = y + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    return y


# ---- functionals.py transliterated (only the north-star terms) ---------------------------------

def thomas_fermi(den, score, Ne):
    return (3. / 10.) * (3. * math.pi ** 2) ** (2 / 3) * (Ne ** (5 / 3)) * den ** (2 / 3)


def thomas_fermi_1D(den, score, Ne):
    return (math.pi * math.pi) / 24 * (Ne ** 3) * den * den


def weizsacker(den, score, Ne, lambda_0=.2):
    return (lambda_0 * Ne / 8.) * torch.einsum('ij,ij->i', score, score)[:, None]


def lda(den, score, Ne):
    l = -(3 / 4) * (Ne ** (4 / 3)) * (3 / math.pi) ** 1 / 3     # functionals.py:415 as coded
    return l * den ** (1 / 3)


def hartree_potential(x, xp, Ne, eps=1E-5):
    z = torch.sum((x - xp) * (x - xp) + eps, dim=-1, keepdim=True)
    return 0.5 * (Ne ** 2) * (1. / (z ** 0.5))


def soft_coulomb(x, xp, Ne):
    return Ne ** 2 / torch.sqrt(1 + (x - xp) * (x - xp))


def nuclei_potential(x, Ne, coords, z, eps=1E-4):
    coords = torch.as_tensor(coords, dtype=DT); z = torch.as_tensor(z, dtype=DT)
    r = torch.stack([torch.sqrt(torch.sum((x - coords[i]) * (x - coords[i]), dim=1)) + eps
                     for i in range(coords.shape[0])], -1)
    return -Ne * torch.sum(z / r, dim=-1, keepdim=True)


def nuclei_potential_hgh(x, Ne, coords, z):
    coords = torch.as_tensor(coords, dtype=DT)
    pp = {'Zion': 1., 'rloc': 0.2, 'C1': -4.180237, 'C2': 0.725075, 'C3': 0., 'C4': 0.}
    tot = 0.
    for i in range(coords.shape[0]):
        r = torch.sqrt(torch.sum((x - coords[i]) * (x - coords[i]), dim=1))
        rr = r / pp['rloc']; rr2 = rr * rr
        v0 = (-pp['Zion'] / r) * torch.erf(rr / math.sqrt(2))
        v1 = torch.exp(-0.5 * rr2)
        v2 = pp['C1'] + pp['C2'] * rr2 + pp['C3'] * rr ** 4 + pp['C4'] * rr ** 6
        tot = tot + v0 + v1 * v2
    return Ne * tot[:, None]


def attraction(x, R, Z_alpha, Z_beta, Ne):
    return Ne * (-Z_alpha / torch.sqrt(1 + (x + R / 2) ** 2) - Z_beta / torch.sqrt(1 + (x - R / 2) ** 2))


def loss(model, p, batch, spec, method="dopri5", n_steps=16, t0=0.0, t1=1.0):
    """H20_...py:150-173 / LiH.py:118-138; ``spec`` is an oracle.ofdft_oracle.EnergySpec."""
    d = model.d
    bs = batch.shape[0] // 2
    fn = lambda y, t: evol_score(model, p, t, y)
    if method == "dopri5":
        y1, stats = odeint_dopri5(fn, batch, t0, t1)
    else:
        y1, stats = odeint_rk4(fn, batch, t0, t1, n_steps), {}
    den_all = torch.exp(y1[:, d:d + 1]); x_all = y1[:, :d]; score_all = y1[:, d + 1:]
    den, x, xp, score = den_all[:bs], x_all[:bs], x_all[bs:], score_all[:bs]
    Ne = spec.Ne
    kin = spec.kin.lower()
    e_t = {"tf": thomas_fermi, "w": weizsacker, "tf1d": thomas_fermi_1D,
           "tf-w": lambda *a: thomas_fermi(*a) + weizsacker(*a),
           "tfw_1d": lambda *a: thomas_fermi_1D(*a) + weizsacker(*a)}[kin](den, score, Ne)
    e_h = {"hartree": hartree_potential, "softc": soft_coulomb}[spec.hart.lower()](x, xp, Ne)
    nn = spec.nuc.lower()
    if nn in ("attraction", "attr"):
        e_n = attraction(x, spec.R, spec.Z_alpha, spec.Z_beta, Ne)
    elif nn in ("hgh", "nuclei_potential_hgh"):
        e_n = nuclei_potential_hgh(x, Ne, spec.coords, spec.z)
    else:
        e_n = nuclei_potential(x, Ne, spec.coords, spec.z)
    e_x = lda(den, score, Ne) if spec.x else torch.zeros_like(e_t)
    e = e_t + e_n + e_h + e_x
    terms = torch.stack([e_t.mean(), e_n.mean(), e_h.mean(), e_x.mean(), torch.zeros((), dtype=DT)])
    return e.mean(), (terms, y1, stats)


def value_and_grad_loss(model, theta_flat, n_in, features, batch, spec, method="dopri5", n_steps=16):
    """jax.value_and_grad(loss, has_aux=True) (H20_...py:177-178) by reverse-mode through the solver."""
    theta = torch.as_tensor(theta_flat, dtype=DT).clone().requires_grad_(True)
    p = params_from_flat(theta, n_in, features, model.d if isinstance(model, GenCNFSimpleMLP) else 1)
    batch = torch.as_tensor(batch, dtype=DT)
    e, (terms, y1, stats) = loss(model, p, batch, spec, method, n_steps)
    g, = torch.autograd.grad(e, theta)
    return e.item(), terms.detach().numpy(), g.detach().numpy(), y1.detach().numpy(), stats
