"""Host-side input producers mirroring ofdft_normflows/utils.py and promolecular_distrax.py.

These are NOT on the accelerated path (the batch generators define the input layout and are
cheap); they exist so a driver written against the reference finds ``coordinates``, ``one_hot_encode``,
``batch_generator`` and ``ProMolecularDensity`` here.  NumPy on the host; the batch is handed to the kernels as a
(2*bs, 2d+1) float32 array.
"""
from __future__ import annotations

import math
from typing import Iterator, Tuple

import numpy as np

BOHR = 1.8897259886


def coordinates(mol_name: str, BOHR: float = BOHR) -> Tuple[int, list, np.ndarray, np.ndarray]:
    """utils.py:276-418: (Ne, atoms, z, coords[bohr]) for H2, LiH, H2O, CH4, C6H6, C27H46O."""
    if mol_name == "H2":
        return 2, ["H", "H"], np.array([1., 1.]), np.array([[0., 0., -1.4008538753 / 2], [0., 0., 1.4008538753 / 2]]) * BOHR
    if mol_name == "LiH":
        return 4, ["Li", "H"], np.array([3., 1.]), np.array([[0., 0., -1.5949 / 2], [0., 0., 1.5949 / 2]]) * BOHR
    if mol_name == "H2O":
        c = np.array([[0.0, 0.0, 0.1189120], [0.0, 0.7612710, -0.4756480], [0.0, -0.7612710, -0.4756480]]) * BOHR
        return 10, ["O", "H", "H"], np.array([8., 1., 1.]), c
    if mol_name == "CH4":
        c = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 1.09], [1.03, 0.0, -0.363], [-0.515, -0.889165, -0.363],
                      [-0.515, 0.889165, -0.363]]) * BOHR
        return 10, ["C", "H", "H", "H", "H"], np.array([6., 1., 1., 1., 1.]), c
    if mol_name == "C6H6":
        c = np.array([[-0.6984022192, 1.2096794375, 0.0001298085], [0.6983971652, 1.2096794375, 0.0001298085],
                      [1.3968148123, 0.0000100819, 0.0001298085], [0.6983970907, -1.2096631450, -0.0001577731],
                      [-0.6984347101, -1.2097236297, -0.0001117213], [-1.3967427862, -0.0000316846, -0.0000531302],
                      [2.4989957176, 0.0000352657, 0.0001433155], [1.2492115488, 2.1643070022, 0.0000608098],
                      [-1.2497352350, 2.1640355443, 0.0000873078], [1.2495192781, -2.1641211865, -0.0004652006],
                      [-1.2494488074, -2.1642720181, -0.0003281198], [-2.4988922798, 0.0006052813, -0.0002941369]]) * BOHR
        return 42, ["C"] * 6 + ["H"] * 6, np.array([6.] * 6 + [1.] * 6), c
    if mol_name == "C27H46O":
        from ._c27h46o import XYZ
        rows = [ln.split() for ln in XYZ.strip().splitlines()]
        atoms = [r[0] for r in rows]
        c = np.array([[float(v) for v in r[1:]] for r in rows]) * BOHR
        return 216, atoms, np.array([{"O": 8., "C": 6., "H": 1.}[a] for a in atoms]), c
    raise KeyError(mol_name)


def one_hot_encode(z) -> np.ndarray:
    """utils.py:251-274 (OFDFT_NF.py:73): one-hot of the index of Z among the sorted unique atomic numbers."""
    z = np.asarray(z)
    zu = np.unique(z)
    out = np.zeros((len(z), len(zu)), dtype=np.float32)
    for i, zi in enumerate(z):
        out[i, int(np.nonzero(zu == zi)[0][0])] = 1.0
    return out


def jax_one_hot(z) -> np.ndarray:
    """``jax.nn.one_hot(z, jnp.unique(z).shape[0]+1)`` as the H2 / H2O / LiH drivers call it
    (H20_mol_ofdft_min_equiv_flows.py:83-84): the class index is the atomic number itself, so any
    Z >= n_types yields an all-zero row (O and Li are encoded as zeros -- reference behaviour)."""
    z = np.asarray(z)
    n = np.unique(z).shape[0] + 1
    out = np.zeros((len(z), n), dtype=np.float32)
    for i, zi in enumerate(z):
        if 0 <= int(zi) < n:
            out[i, int(zi)] = 1.0
    return out


class ProMolecularDensity:
    """promolecular_distrax.py:18-89: mixture of unit-variance Gaussians on the nuclei, weights Z_i / sum Z."""

    def __init__(self, z, loc):
        self.z = np.asarray(z, dtype=np.float64).ravel()
        self.loc = np.asarray(loc, dtype=np.float64).reshape(-1, 3)
        self.probs = self.z / np.abs(self.z).sum()

    def sample(self, rng: np.random.Generator, n: int) -> np.ndarray:
        comp = rng.choice(len(self.z), size=n, p=self.probs)
        return self.loc[comp] + rng.standard_normal((n, 3))

    def log_prob_and_score(self, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        dv = x[:, None, :] - self.loc[None]
        lg = -0.5 * (dv * dv).sum(-1) - 1.5 * math.log(2.0 * math.pi)
        m = lg.max(1, keepdims=True)
        pe = np.exp(lg - m) * self.probs[None]
        tot = pe.sum(1, keepdims=True)
        return np.log(tot) + m, -((pe / tot)[..., None] * dv).sum(1)

    def log_prob(self, x: np.ndarray) -> np.ndarray:
        return self.log_prob_and_score(x)[0]

    def score(self, x: np.ndarray) -> np.ndarray:
        return self.log_prob_and_score(x)[1]


def batch_generator(seed: int, batch_size: int, prior_dist: ProMolecularDensity) -> Iterator[np.ndarray]:
    """utils.py:76-108: yields (2*bs, 7) float32 rows [x, log rho0(x), grad log rho0(x)] -- two independent draws."""
    rng = np.random.default_rng(seed)
    while True:
        halves = []
        for _ in range(2):
            x = prior_dist.sample(rng, batch_size)
            lp, sc = prior_dist.log_prob_and_score(x)
            halves.append(np.concatenate([x, lp, sc], 1))
        yield np.concatenate(halves, 0).astype(np.float32)


def batche_generator_1D(seed: int, batch_size: int) -> Iterator[np.ndarray]:
    """utils.py:110-143 with the standard-normal prior of LiH.py:78: (2*bs, 3) rows [x, logp0, score0]."""
    rng = np.random.default_rng(seed)
    while True:
        halves = []
        for _ in range(2):
            x = rng.standard_normal((batch_size, 1))
            halves.append(np.concatenate([x, -0.5 * x * x - 0.5 * math.log(2 * math.pi), -x], 1))
        yield np.concatenate(halves, 0).astype(np.float32)


# ---- device-side variants (CUDA kernels of csrc/prior.cu) -----------------------------

def _prior_tables(prior: ProMolecularDensity):
    coords = np.ascontiguousarray(prior.loc, dtype=np.float32)
    z = np.ascontiguousarray(prior.z, dtype=np.float32)
    return coords, z


def promolecular_log_prob_score_device(prior: ProMolecularDensity, x, want_score: bool = True):
    """log rho0(x) (n,1) and grad log rho0(x) (n,3) for a CUDA tensor x (n, >=3) -- ofdft_promolecular_logp_score."""
    import torch
    from . import _lib, ops
    lib = _lib.load()
    x = ops._f32c(x)
    n = x.shape[0]
    coords, z = _prior_tables(prior)
    out = torch.empty((n, 1), dtype=torch.float32, device=x.device)
    score = torch.empty((n, 3), dtype=torch.float32, device=x.device) if want_score else None
    with ops._on(x):
        _lib.check(lib.ofdft_promolecular_logp_score(ops._stream(x), coords.ctypes.data, z.ctypes.data, len(z), ops._ptr(x),
                                                     x.shape[1], n, None, 0, 0, ops._ptr(out), ops._ptr(score)),
                   "ofdft_promolecular_logp_score")
    return out, score


def batch_generator_device(seed: int, batch_size: int, prior: ProMolecularDensity, device="cuda"):
    """utils.py:76-108 on the device: yields (2*bs, 7) CUDA tensors [x, log rho0, grad log rho0] (Philox, reproducible
    per (seed, call index)); no host input is needed for a training step."""
    import torch
    from . import _lib, ops
    lib = _lib.load()
    coords, z = _prior_tables(prior)
    call = 0
    while True:
        out = torch.empty((2 * batch_size, 7), dtype=torch.float32, device=device)
        with torch.cuda.device(out.device):
            _lib.check(lib.ofdft_batch_generator_3d(ops._stream(out), seed, call, coords.ctypes.data, z.ctypes.data, len(z),
                                                    batch_size, ops._ptr(out)), "ofdft_batch_generator_3d")
        call += 1
        yield out


def batche_generator_1D_device(seed: int, batch_size: int, device="cuda"):
    """utils.py:110-143 on the device (N(0,1) prior of LiH.py:78): yields (2*bs, 3) CUDA tensors."""
    import torch
    from . import _lib, ops
    lib = _lib.load()
    call = 0
    while True:
        out = torch.empty((2 * batch_size, 3), dtype=torch.float32, device=device)
        with torch.cuda.device(out.device):
            _lib.check(lib.ofdft_batch_generator_1d(ops._stream(out), seed, call, batch_size, ops._ptr(out)), "ofdft_batch_generator_1d")
        call += 1
        yield out


def rho_rev(params, x, model_rev, prior: ProMolecularDensity, n_steps: int = None):
    """rho_rev(params, x) of the 3-D drivers (H20_mol_ofdft_min_equiv_flows.py:132-137): density of the flow at the points
    x (n,3): z_t = [x, 0] -> neural_ode(model_rev, -1 -> 0) -> exp(log rho0(z0) - Delta logp).  Returns (n,1)."""
    import torch
    from . import _lib, ops
    from .jax_ode import DEFAULT_N_STEPS
    lib = _lib.load()
    x = ops._f32c(x)
    n = x.shape[0]
    zt = torch.zeros((n, 4), dtype=torch.float32, device=x.device)
    zt[:, :3] = x[:, :3]
    theta = model_rev.flat_theta(params).detach()
    y0, _ = ops.gnn_fwd(theta, model_rev.xyz_nuclei, model_rev.z_one_hot, zt, n, ops.JETS_XLOGP, ops.JETS_XLOGP,
                        n_steps or DEFAULT_N_STEPS, -1.0, 0.0, model_rev.y0)
    coords, z = _prior_tables(prior)
    rho = torch.empty((n, 1), dtype=torch.float32, device=x.device)
    with ops._on(y0):
        _lib.check(lib.ofdft_promolecular_logp_score(ops._stream(y0), coords.ctypes.data, z.ctypes.data, len(z), ops._ptr(y0), 4, n,
                                                     y0.data_ptr() + 12, 4, 1, ops._ptr(rho), None), "ofdft_promolecular_logp_score")
    return rho


def rho_rev_1d(params, x, model_rev, n_steps: int = None):
    """rho_rev(params, x) of the 1-D driver (LiH.py:104-109): density of the flow at the points x (n,1) under the N(0,1) prior
    (LiH.py:78): z_t = [x, 0] -> neural_ode(model_rev, -1 -> 0) -> exp(log N(z0) - Delta logp).  Returns (n,1)."""
    import math
    import torch
    from . import ops
    from .cn_flows import mlp_fwd
    from .jax_ode import DEFAULT_N_STEPS
    x = ops._f32c(x)
    n = x.shape[0]
    zt = torch.zeros((n, 2), dtype=torch.float32, device=x.device)
    zt[:, 0] = x[:, 0]
    theta = model_rev.flat_theta(params).detach()
    y0, _ = mlp_fwd(theta, zt, model_rev, -1.0, 0.0, ops.JETS_XLOGP, n_steps or DEFAULT_N_STEPS)
    # the base log-density is three flops per point: plain tensor arithmetic on the device result
    logp0 = -0.5 * y0[:, 0:1] ** 2 - 0.5 * math.log(2.0 * math.pi)
    return torch.exp(logp0 - y0[:, 1:2])


def compute_integral(params, grid_array, rho, Ne: float, bs: int = 0):
    """compute_integral (H20_mol_ofdft_min_equiv_flows.py:35-39): Ne * <grid_weights, rho(params, grid_coords)>; ``rho`` is a
    callable such as ``lambda p, x: rho_rev(p, x, model_rev, prior)``."""
    import torch
    grid_coords, grid_weights = grid_array
    rho_val = Ne * rho(params, grid_coords)
    w = torch.as_tensor(grid_weights, dtype=rho_val.dtype, device=rho_val.device).reshape(-1)
    return torch.dot(w, rho_val.reshape(-1))
