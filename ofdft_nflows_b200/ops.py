"""Device-level operators: thin wrappers that pass torch CUDA buffers and the current stream to the C ABI.

torch is used for device memory, streams and (elsewhere) torch.distributed only; every arithmetic
operation of the hot path happens inside libofdft_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import EnergySpecC

JETS_X, JETS_XLOGP, JETS_FULL = 1, 2, 3
# kernel-path selector of the integrator entry points (include/ofdft_b200.h): the default runs the GEMM-shaped phases on
# the tcgen05 tensor cores; PATH_FFMA / PATH_FFMA_DW select the independently written FP32-pipe kernels
PATH_DEFAULT, PATH_FFMA, PATH_FFMA_DW = 0, 1, 2
PATH_DEEP = 8           # 3-D flow, (64, 64): depth-generic tensor-core kernels (the default of (64, 64, 64)) as a cross-check
PATH_ACT_CKPT = 4       # 1-D flow: the checkpoint buffer also holds the layer activations (set by cn_flows.mlp_fwd / mlp_bwd)

KIN = {"": 0, "none": 0, "tf": 1, "thomas_fermi": 1, "w": 2, "weizsacker": 2, "w1d": 2, "weizsacker1d": 2,
       "tf-w": 3, "thomas_fermi_weizsacker": 3, "tf1d": 4, "thomas_fermi_1d": 4,
       "tfw_1d": 5, "thomas_fermi_weizsacker_1d": 5}
HART = {"": 0, "none": 0, "hartree": 1, "hartree_potential": 1, "softc": 2, "soft_coulomb": 2, "hartree_mt": 3}
NUC = {"": 0, "none": 0, "nuclei_potential": 1, "nuclei": 1, "hgh": 2, "nuclei_potential_hgh": 2,
       "attraction": 3, "attr": 3, "hgh_species": 4}
XC_X = {"": 0, "none": 0, "lda": 1, "local_density_approximation": 1, "b88_x_e": 2, "dirac_b88_x_e": 3}
XC_C = {"": 0, "none": 0, "vwn_c_e": 1, "correlation_vwn_c_e": 1, "pw92_c_e": 2, "correlation_pw92_c_e": 2,
        "xc_1d": 3, "exchange_correlation_1d": 3}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(t: Optional[torch.Tensor] = None) -> int:
    return torch.cuda.current_stream(None if t is None else t.device).cuda_stream


def _on(t: torch.Tensor):
    """The C launchers query the current device: make it the tensors' device for the duration of the call."""
    return torch.cuda.device(t.device)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("ofdft_nflows_b200: tensors must live on a CUDA device (no CPU fallback)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


def _cols(jets: int, d: int = 3) -> int:
    return {1: d, 2: d + 1, 3: 2 * d + 1}[jets]


def gnn_theta_size(n_types: int, n_layers: int = 2) -> int:
    return 1 + 64 + 64 + 64 * (2 + n_types) + (n_layers - 1) * (64 + 64 * 64)


def gnn_n_layers(theta: torch.Tensor, n_types: int) -> int:
    for L in (1, 2, 3):
        if theta.numel() == gnn_theta_size(n_types, L):
            return L
    raise ValueError(f"theta has {theta.numel()} entries: not a Gen_EqvFlow stack of 1-3 hidden layers of 64 units "
                     f"with {n_types} species")


def gnn_act_ckpt_bytes(B: int, n_a: int, jets_a: int, jets_b: int, Na: int, n_steps: int) -> int:
    return int(_lib.load().ofdft_cnf_gnn_act_ckpt_bytes(B, n_a, jets_a, jets_b, Na, n_steps))


def gnn_fwd(theta, nuclei, onehot, y_in, n_a: int, jets_a: int, jets_b: int, n_steps: int,
            t0: float, t1: float, y0sign: float, want_ckpt: bool = False, out: Optional[torch.Tensor] = None,
            want_act: bool = False, path: int = PATH_DEFAULT):
    """ofdft_cnf_gnn_fwd: integrate rows [0,n_a) with jets_a and rows [n_a,B) with jets_b.
    Returns (y1, ckpt) -- ckpt is None, the stage-state checkpoint, or (with ``want_act``) the pair
    (stage states, layer-2 activation checkpoint)."""
    lib = _lib.load()
    theta, nuclei, onehot, y_in = _f32c(theta), _f32c(nuclei), _f32c(onehot), _f32c(y_in)
    B, stride = y_in.shape
    Na, n_types = onehot.shape
    L = gnn_n_layers(theta, n_types)
    y_out = out if out is not None else torch.zeros_like(y_in)
    ckpt = None
    if want_ckpt:
        ckpt = torch.empty(lib.ofdft_cnf_gnn_ckpt_bytes(B, n_steps) // 4, dtype=torch.float32, device=y_in.device)
    act = None
    if want_ckpt and want_act and L == 2 and not (int(path) & PATH_DEEP):      # the depth-generic kernels keep no activation checkpoint
        nb = lib.ofdft_cnf_gnn_act_ckpt_bytes(B, n_a, jets_a, jets_b, Na, n_steps)
        act = torch.empty(nb // 4, dtype=torch.float32, device=y_in.device)
    sb = lib.ofdft_cnf_gnn_fwd_scratch_bytes()
    scratch = torch.empty(sb, dtype=torch.uint8, device=y_in.device)
    with _on(y_in):
        _lib.check(lib.ofdft_cnf_gnn_fwd(_stream(y_in), _ptr(theta), _ptr(nuclei), _ptr(onehot), _ptr(y_in), _ptr(y_out),
                                         _ptr(ckpt), _ptr(act), _ptr(scratch), sb, B, n_a, jets_a, jets_b, stride,
                                         y_out.shape[1], Na, n_types, L, n_steps, t0, t1, y0sign, int(path)), "ofdft_cnf_gnn_fwd")
    return y_out, ((ckpt, act) if act is not None else ckpt)


def gnn_bwd(theta, nuclei, onehot, ckpt, ybar1, n_a: int, jets_a: int, jets_b: int, n_steps: int,
            t0: float, t1: float, y0sign: float, want_ybar0: bool = False, path: int = PATH_DEFAULT):
    """ofdft_cnf_gnn_bwd: discrete adjoint; returns (theta_bar, ybar0 or None)."""
    lib = _lib.load()
    theta, nuclei, onehot, ybar1 = _f32c(theta), _f32c(nuclei), _f32c(onehot), _f32c(ybar1)
    act = None
    if isinstance(ckpt, tuple):
        ckpt, act = ckpt
    B, stride = ybar1.shape
    Na, n_types = onehot.shape
    theta_bar = torch.empty_like(theta)
    ybar0 = torch.zeros_like(ybar1) if want_ybar0 else None
    L = gnn_n_layers(theta, n_types)
    sb = lib.ofdft_cnf_gnn_bwd_scratch_bytes(B, n_a, n_types, L)
    scratch = torch.empty(sb, dtype=torch.uint8, device=ybar1.device)
    with _on(ybar1):
        _lib.check(lib.ofdft_cnf_gnn_bwd(_stream(ybar1), _ptr(theta), _ptr(nuclei), _ptr(onehot), _ptr(ckpt), _ptr(act),
                                         _ptr(ybar1), _ptr(theta_bar), _ptr(ybar0), _ptr(scratch), sb, B, n_a, jets_a, jets_b,
                                         stride, Na, n_types, L, n_steps, t0, t1, y0sign, int(path)), "ofdft_cnf_gnn_bwd")
    return theta_bar, ybar0


def gnn_rhs(theta, nuclei, onehot, y, jets: int, t: float, y0sign: float, path: int = PATH_DEFAULT) -> torch.Tensor:
    """ofdft_cnf_gnn_rhs: one evaluation of d/dt [x, logp, (score)]."""
    lib = _lib.load()
    theta, nuclei, onehot, y = _f32c(theta), _f32c(nuclei), _f32c(onehot), _f32c(y)
    B, stride = y.shape
    Na, n_types = onehot.shape
    dy = torch.zeros_like(y)
    sb = lib.ofdft_cnf_gnn_fwd_scratch_bytes()
    scratch = torch.empty(sb, dtype=torch.uint8, device=y.device)
    with _on(y):
        _lib.check(lib.ofdft_cnf_gnn_rhs(_stream(y), _ptr(theta), _ptr(nuclei), _ptr(onehot), _ptr(y), _ptr(dy), _ptr(scratch),
                                         sb, B, jets, stride, Na, n_types, gnn_n_layers(theta, n_types), t, y0sign, int(path)),
                   "ofdft_cnf_gnn_rhs")
    return dy


@dataclass
class EnergySpec:
    """Functional selection of one driver run (argparse names of the reference drivers,
    H20_mol_ofdft_min_equiv_flows.py:225-249, LiH.py:211-223)."""
    Ne: float
    d: int = 3
    kin: str = "tf-w"
    hart: str = "hartree"
    nuc: str = "nuclei_potential"
    x: str = "lda"
    c: str = ""                       # c-type functional (den, Ne): vwn_c_e | pw92_c_e | xc_1d
    coords: Optional[object] = None   # (Na,3) host array (bohr)
    z: Optional[object] = None        # (Na,)
    R: float = 0.0
    Z_alpha: float = 0.0
    Z_beta: float = 0.0

    def to_c(self):
        import numpy as np
        try:
            kin, hart, nuc, x = KIN[self.kin.lower()], HART[self.hart.lower()], NUC[self.nuc.lower()], XC_X[self.x.lower()]
            cc = XC_C[self.c.lower()]
        except KeyError as e:
            raise NotImplementedError(f"functional {e} is not implemented by the CUDA path") from None
        n = 0
        coords = z = None
        if self.d == 3 and nuc != 0:
            coords = np.ascontiguousarray(np.asarray(self.coords, dtype=np.float32).reshape(-1, 3))
            z = np.ascontiguousarray(np.asarray(self.z, dtype=np.float32).reshape(-1))
            n = coords.shape[0]
        spec = EnergySpecC(self.d, kin, hart, nuc, x, cc, n, float(self.Ne), float(self.R), float(self.Z_alpha), float(self.Z_beta))
        return spec, coords, z


def _host_ptr(a):
    return None if a is None else a.ctypes.data


def functionals_fwd_bwd(spec: EnergySpec, y1: torch.Tensor, want_bar: bool = True):
    """ofdft_functionals_fwd_bwd: returns (sums float64[8] on device: means kin, vnuc, hart, x, c, total; ybar1)."""
    lib = _lib.load()
    y1 = _f32c(y1)
    B, stride = y1.shape
    bs = B // 2
    cs, coords, z = spec.to_c()
    sums = torch.empty(8, dtype=torch.float64, device=y1.device)
    ybar1 = torch.empty_like(y1) if want_bar else None
    sb = lib.ofdft_functionals_scratch_bytes(bs)
    scratch = torch.empty(sb, dtype=torch.uint8, device=y1.device)
    with _on(y1):
        _lib.check(lib.ofdft_functionals_fwd_bwd(_stream(y1), C.byref(cs), _host_ptr(coords), _host_ptr(z), _ptr(y1), stride,
                                                 bs, _ptr(sums), _ptr(ybar1), _ptr(scratch), sb), "ofdft_functionals_fwd_bwd")
    return sums, ybar1


def functionals_integrands(spec: EnergySpec, y1: torch.Tensor) -> torch.Tensor:
    """ofdft_functionals_integrands: (5, bs) per-sample integrands [kin, vnuc, hart, x, c]."""
    lib = _lib.load()
    y1 = _f32c(y1)
    B, stride = y1.shape
    bs = B // 2
    cs, coords, z = spec.to_c()
    e = torch.empty((5, bs), dtype=torch.float32, device=y1.device)
    with _on(y1):
        _lib.check(lib.ofdft_functionals_integrands(_stream(y1), C.byref(cs), _host_ptr(coords), _host_ptr(z), _ptr(y1), stride,
                                                    bs, _ptr(e)), "ofdft_functionals_integrands")
    return e


def hartree_allpairs(x: torch.Tensor, xp: torch.Tensor, Ne: float, kind: str = "hartree", want_grad: bool = True,
                     xbar_into: Optional[torch.Tensor] = None, xpbar_into: Optional[torch.Tensor] = None,
                     norm_pairs: float = 0.0, esum_into: Optional[torch.Tensor] = None):
    """ofdft_hartree_allpairs: cross-set pair mean + gradients.  x (n, >=d), xp (m, >=d): state rows or packed coordinates
    (the row strides may differ).  ``xbar_into`` / ``xpbar_into``: contiguous (n|m, stride) cotangent buffers whose first d
    columns are ADDED to (the kernel accumulates in place, and into ``esum_into`` when given); otherwise fresh (n|m, d)
    tensors are returned.  ``norm_pairs``: size of the pair set the mean is taken over (default n*m)."""
    lib = _lib.load()
    x, xp = _f32c(x), _f32c(xp)
    d = 3 if HART[kind] == 1 else 1
    n, m = x.shape[0], xp.shape[0]
    accumulate = xbar_into is not None or xpbar_into is not None or esum_into is not None
    esum = esum_into if esum_into is not None else torch.zeros(1, dtype=torch.float64, device=x.device)
    if accumulate:
        xbar, xpbar = xbar_into, xpbar_into
        if any(not t.is_contiguous() or t.dtype != torch.float32 for t in (xbar, xpbar) if t is not None):
            raise ValueError("xbar_into / xpbar_into must be contiguous float32 buffers")
    else:
        xbar = torch.empty((n, d), dtype=torch.float32, device=x.device) if want_grad else None
        xpbar = torch.empty((m, d), dtype=torch.float32, device=x.device) if want_grad else None
    sb = lib.ofdft_hartree_allpairs_scratch_bytes(n, m)
    scratch = torch.empty(sb, dtype=torch.uint8, device=x.device)
    with _on(x):
        _lib.check(lib.ofdft_hartree_allpairs(_stream(x), HART[kind], d, float(Ne), _ptr(x), x.shape[1], _ptr(xp), xp.shape[1],
                                              n, m, float(norm_pairs), _ptr(esum), _ptr(xbar), 0 if xbar is None else xbar.shape[1],
                                              _ptr(xpbar), 0 if xpbar is None else xpbar.shape[1], int(accumulate),
                                              _ptr(scratch), sb), "ofdft_hartree_allpairs")
    return esum, xbar, xpbar


def microbench(kind: str, iters: int = 2000, repeat: int = 5) -> float:
    """Measured FP32-FFMA (flop/s, 2 per FFMA) or SFU (rsqrt/s) peak of the current device."""
    lib = _lib.load()
    sink = torch.zeros(4, dtype=torch.float32, device="cuda")
    ops = C.c_double(0.0)
    fn = lib.ofdft_microbench_ffma if kind == "ffma" else lib.ofdft_microbench_sfu
    best = float("inf")
    for _ in range(repeat + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(fn(_stream(), _ptr(sink), iters, C.byref(ops)), "microbench")
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return ops.value * (2.0 if kind == "ffma" else 1.0) / best


def rmsprop_clip_step(theta: torch.Tensor, grad: torch.Tensor, nu: torch.Tensor, lr: float, decay: float = 0.9,
                      eps: float = 1e-8, max_norm: float = 1.0) -> None:
    """ofdft_rmsprop_clip_step: optax.chain(clip_by_global_norm(max_norm), rmsprop(lr)) in place on flat theta / nu."""
    lib = _lib.load()
    for t in (theta, grad, nu):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("theta / grad / nu must be contiguous float32 CUDA tensors")
    with _on(theta):
        _lib.check(lib.ofdft_rmsprop_clip_step(_stream(theta), _ptr(theta), _ptr(grad), _ptr(nu), theta.numel(), float(lr),
                                               float(decay), float(eps), float(max_norm)), "ofdft_rmsprop_clip_step")


def adam_clip_step(theta: torch.Tensor, grad: torch.Tensor, mu: torch.Tensor, nu: torch.Tensor, step: int, lr: float,
                   b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, max_norm: float = 1.0) -> None:
    """ofdft_adam_clip_step: optax.chain(clip_by_global_norm(max_norm), adam(lr)) (OFDFT_NF.py:104-108) in place on the flat
    theta and the moment states mu / nu; ``step`` is the 1-based update count."""
    lib = _lib.load()
    for t in (theta, grad, mu, nu):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("theta / grad / mu / nu must be contiguous float32 CUDA tensors")
    with _on(theta):
        _lib.check(lib.ofdft_adam_clip_step(_stream(theta), _ptr(theta), _ptr(grad), _ptr(mu), _ptr(nu), theta.numel(), int(step),
                                            float(lr), float(b1), float(b2), float(eps), float(max_norm)), "ofdft_adam_clip_step")


class EnergyEMA:
    """optax.ema(decay, debias=True) of the per-term energy means as the drivers log it
    (H20_mol_ofdft_min_equiv_flows.py:117-118, 193-196), device-resident: ``update(sums)`` takes the float64 device
    vector returned by the fused step and returns the debiased averages; ``history`` keeps every update on the device so
    a training loop reads it back once at the end instead of once per epoch."""

    def __init__(self, n: int = 6, decay: float = 0.99, capacity: int = 0, device="cuda"):
        self.decay, self.n, self.step = float(decay), int(n), 0
        self.state = torch.zeros(n, dtype=torch.float64, device=device)
        self.history = torch.zeros((max(capacity, 1), n), dtype=torch.float64, device=device)
        self.capacity = int(capacity)

    def update(self, values: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        if not values.is_cuda or values.dtype != torch.float64 or values.numel() != self.n:
            raise ValueError("values must be a float64 CUDA vector of length n")
        self.step += 1
        row = (self.step - 1) % self.history.shape[0] if self.capacity else 0
        out = self.history[row]
        with _on(values):
            _lib.check(lib.ofdft_ema_update(_stream(values), _ptr(values.contiguous()), _ptr(self.state), _ptr(out), self.n,
                                            self.step, self.decay), "ofdft_ema_update")
        return out
