"""jax.ffi + jax.custom_vjp drop-ins for ofdft_normflows over the C ABI (include/ofdft_b200.h, ABI version 3).

Import-guarded: JAX is NOT installed in the build image of this repository, so this module has been checked there only
statically (tests/test_capi.py: it parses, every FFI target it registers is defined by xla_ffi_shim.cc, every size helper it
calls is exported by libofdft_b200.so); it runs where JAX >= 0.4.38 and libofdft_b200_xla.so (xla_ffi_shim.cc) exist.

    from ofdft_nflows_b200.jax_ffi.bindings import neural_ode, neural_ode_score     # same signatures as jax_ode.py:9,51
    from ofdft_nflows_b200.jax_ffi.bindings import energy_terms, hartree_allpairs   # fused functionals / pair kernel

Every op is a ``jax.ffi.ffi_call`` whose extra results are the buffers the kernels need (stage / activation checkpoints,
scratch), sized by the ``*_bytes`` helpers of the C ABI, wrapped in ``jax.custom_vjp`` so that ``jax.value_and_grad(loss)``
(H20_mol_ofdft_min_equiv_flows.py:177-178, LiH.py:145-146) reaches the discrete-adjoint kernels.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

try:
    import jax
    import jax.numpy as jnp
    import numpy as np
    from jax.flatten_util import ravel_pytree
except ImportError as e:  # pragma: no cover - JAX absent in this image
    raise ImportError("ofdft_nflows_b200.jax_ffi needs JAX (jax.ffi); it is not installed in this environment") from e

_HERE = Path(__file__).resolve().parent.parent
_xla = ctypes.CDLL(str(_HERE / "libofdft_b200_xla.so"))
_abi = ctypes.CDLL(str(_HERE / "libofdft_b200.so"))
FFI_TARGETS = ("OfdftGnnFwd", "OfdftGnnBwd", "OfdftGnnRhs", "OfdftMlp1dFwd", "OfdftMlp1dBwd", "OfdftMlp1dRhs",
               "OfdftFunctionals", "OfdftHartreeAllpairs")
for _name in FFI_TARGETS:
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_xla, _name)), platform="CUDA")

SIZE_HELPERS = ("ofdft_cnf_gnn_ckpt_bytes", "ofdft_cnf_gnn_act_ckpt_bytes", "ofdft_cnf_gnn_fwd_scratch_bytes",
                "ofdft_cnf_gnn_bwd_scratch_bytes", "ofdft_cnf_mlp1d_ckpt_bytes", "ofdft_cnf_mlp1d_act_ckpt_bytes",
                "ofdft_cnf_mlp1d_scratch_bytes",
                "ofdft_functionals_scratch_bytes", "ofdft_hartree_allpairs_scratch_bytes")
for _name in SIZE_HELPERS:
    getattr(_abi, _name).restype = ctypes.c_size_t
_abi.ofdft_cnf_gnn_theta_size.restype = ctypes.c_int64
_i64, _i32 = ctypes.c_int64, ctypes.c_int32
PATH_DEFAULT, PATH_FFMA, PATH_FFMA_DW, PATH_ACT_CKPT = 0, 1, 2, 4
MLP1D_ACT_CKPT_MAX_BYTES = 16 << 30

KIN = {"": 0, "tf": 1, "w": 2, "tf-w": 3, "tf1d": 4, "tfw_1d": 5}
HART = {"": 0, "hartree": 1, "softc": 2, "hartree_mt": 3}
NUC = {"": 0, "nuclei_potential": 1, "hgh": 2, "attraction": 3, "attr": 3}
XC_X = {"": 0, "lda": 1, "b88_x_e": 2, "dirac_b88_x_e": 3}
XC_C = {"": 0, "vwn_c_e": 1, "pw92_c_e": 2, "xc_1d": 3}


def _f32(n):
    return jax.ShapeDtypeStruct((int(n),), jnp.float32)


def _u8(n):
    return jax.ShapeDtypeStruct((int(n),), jnp.uint8)


def _gnn_layers(theta, n_types):
    for L in (1, 2, 3):
        if theta.shape[0] == _abi.ofdft_cnf_gnn_theta_size(_i32(n_types), _i32(L)):
            return L
    raise ValueError("theta is not a Gen_EqvFlow stack of 1-3 hidden layers of 64 units")


# ---- 3-D equivariant GNN flow ---------------------------------------------------------------------------------------------
def make_gnn_flow(nuclei, onehot, *, bool_neg, jets, n_steps=16, t0=0.0, t1=1.0, act_ckpt=True, path=PATH_DEFAULT):
    """Returns ``flow(theta_flat, y0) -> y1`` with a custom VJP (discrete adjoint kernel)."""
    y0sign = -1.0 if bool_neg else 1.0
    nuclei = jnp.asarray(nuclei, jnp.float32)
    onehot = jnp.asarray(onehot, jnp.float32)
    Na, n_types = onehot.shape
    attrs = dict(jets_a=np.int32(jets), jets_b=np.int32(jets), n_steps=np.int32(n_steps), t0=np.float32(t0), t1=np.float32(t1),
                 y0sign=np.float32(y0sign), path=np.int32(path))

    def call_fwd(theta, y0, want_ckpt):
        B, w = y0.shape
        L = _gnn_layers(theta, n_types)
        ckpt_n = _abi.ofdft_cnf_gnn_ckpt_bytes(_i64(B), _i32(n_steps)) // 4 if want_ckpt else 1
        act_n = (_abi.ofdft_cnf_gnn_act_ckpt_bytes(_i64(B), _i64(B), _i32(jets), _i32(jets), _i32(Na), _i32(n_steps)) // 4
                 if (want_ckpt and act_ckpt and L == 2) else 1)
        out = (jax.ShapeDtypeStruct((B, w), jnp.float32), _f32(ckpt_n), _f32(act_n), _u8(_abi.ofdft_cnf_gnn_fwd_scratch_bytes()))
        y1, ckpt, act, _ = jax.ffi.ffi_call("OfdftGnnFwd", out)(theta, nuclei, onehot, y0, n_a=np.int64(B), n_layers=np.int32(L),
                                                                **attrs)
        return y1, ckpt, act

    @jax.custom_vjp
    def flow(theta, y0):
        return call_fwd(theta, y0, False)[0]

    def fwd(theta, y0):
        y1, ckpt, act = call_fwd(theta, y0, True)
        return y1, (theta, ckpt, act)

    def bwd(res, ybar1):
        theta, ckpt, act = res
        B, w = ybar1.shape
        L = _gnn_layers(theta, n_types)
        sb = _abi.ofdft_cnf_gnn_bwd_scratch_bytes(_i64(B), _i64(B), _i32(n_types), _i32(L))
        out = (jax.ShapeDtypeStruct(theta.shape, jnp.float32), jax.ShapeDtypeStruct((B, w), jnp.float32), _u8(sb))
        tbar, ybar0, _ = jax.ffi.ffi_call("OfdftGnnBwd", out)(theta, nuclei, onehot, ckpt, act, ybar1, n_a=np.int64(B),
                                                              n_layers=np.int32(L), **attrs)
        return tbar, ybar0

    flow.defvjp(fwd, bwd)
    return flow


def gnn_rhs(theta, nuclei, onehot, y, t, *, bool_neg, jets, path=PATH_DEFAULT):
    """``f.apply(params, t, states)`` (equiv_flows.py:124-127) / the score-augmented right-hand side (jax_ode.py:80-94)."""
    nuclei = jnp.asarray(nuclei, jnp.float32); onehot = jnp.asarray(onehot, jnp.float32)
    out = (jax.ShapeDtypeStruct(y.shape, jnp.float32), _u8(_abi.ofdft_cnf_gnn_fwd_scratch_bytes()))
    return jax.ffi.ffi_call("OfdftGnnRhs", out)(theta, nuclei, onehot, y, jets=np.int32(jets),
                                                n_layers=np.int32(_gnn_layers(theta, onehot.shape[1])), t=np.float32(t),
                                                y0sign=np.float32(-1.0 if bool_neg else 1.0), path=np.int32(path))[0]


# ---- 1-D tanh-MLP flow ------------------------------------------------------------------------------------------------------
def make_mlp1d_flow(features, *, bool_neg, n_steps=16, t0=0.0, t1=1.0, path=PATH_DEFAULT, act_ckpt=True):
    """``flow(theta_flat, y0 (B,3)) -> y1`` for Gen_CNFSimpleMLP(1, features) (cn_flows.py:166-182), rows [x, logp, score];
    the adjoint kernels carry all three channels, so ``neural_ode`` pads a zero score column (see below).  ``act_ckpt``: the
    residual also keeps the layer activations of every evaluation (OFDFT_PATH_ACT_CKPT) when the shape supports it."""
    H, L = int(features[0]), len(features)
    y0sign = -1.0 if bool_neg else 1.0
    base = dict(jets=np.int32(3), H=np.int32(H), n_layers=np.int32(L), n_steps=np.int32(n_steps), t0=np.float32(t0),
                t1=np.float32(t1), y0sign=np.float32(y0sign))

    def act_bytes(B):
        """Size of the activation checkpoint, or 0 when this batch / path does not use it (same rule on both sides of the vjp)."""
        if not act_ckpt or (path & (PATH_FFMA | PATH_FFMA_DW)):
            return 0
        n = _abi.ofdft_cnf_mlp1d_act_ckpt_bytes(_i64(B), _i32(H), _i32(L), _i32(n_steps))
        return n if 0 < n <= MLP1D_ACT_CKPT_MAX_BYTES else 0

    def call_fwd(theta, y0, want_ckpt):
        B = y0.shape[0]
        n_act = act_bytes(B) if want_ckpt else 0
        ckpt_n = (n_act or _abi.ofdft_cnf_mlp1d_ckpt_bytes(_i64(B), _i32(n_steps))) // 4 if want_ckpt else 1
        sb = _abi.ofdft_cnf_mlp1d_scratch_bytes(_i64(B), _i32(H), _i32(L), _i32(n_steps))
        out = (jax.ShapeDtypeStruct(y0.shape, jnp.float32), _f32(ckpt_n), _u8(sb))
        attrs = dict(base, path=np.int32(path | (PATH_ACT_CKPT if n_act else 0)))
        y1, ckpt, _ = jax.ffi.ffi_call("OfdftMlp1dFwd", out)(theta, y0, **attrs)
        return y1, ckpt

    @jax.custom_vjp
    def flow(theta, y0):
        return call_fwd(theta, y0, False)[0]

    def fwd(theta, y0):
        y1, ckpt = call_fwd(theta, y0, True)
        return y1, (theta, ckpt)

    def bwd(res, ybar1):
        theta, ckpt = res
        B = ybar1.shape[0]
        sb = _abi.ofdft_cnf_mlp1d_scratch_bytes(_i64(B), _i32(H), _i32(L), _i32(n_steps))
        out = (jax.ShapeDtypeStruct(theta.shape, jnp.float32), jax.ShapeDtypeStruct(ybar1.shape, jnp.float32), _u8(sb))
        attrs = dict(base, path=np.int32(path | (PATH_ACT_CKPT if act_bytes(B) else 0)))
        tbar, ybar0, _ = jax.ffi.ffi_call("OfdftMlp1dBwd", out)(theta, ckpt, ybar1, **attrs)
        return tbar, ybar0

    flow.defvjp(fwd, bwd)
    return flow


def mlp1d_rhs(theta, y, t, features, *, bool_neg):
    H, L = int(features[0]), len(features)
    sb = _abi.ofdft_cnf_mlp1d_scratch_bytes(_i64(y.shape[0]), _i32(H), _i32(L), _i32(1))
    out = (jax.ShapeDtypeStruct(y.shape, jnp.float32), _u8(sb))
    return jax.ffi.ffi_call("OfdftMlp1dRhs", out)(theta, y, jets=np.int32({1: 1, 2: 2, 3: 3}[y.shape[1]]), H=np.int32(H),
                                                  n_layers=np.int32(L), t=np.float32(t),
                                                  y0sign=np.float32(-1.0 if bool_neg else 1.0))[0]


# ---- reference entry points -------------------------------------------------------------------------------------------------
def _is_gnn(f):
    return hasattr(f, "xyz_nuclei")


def neural_ode_score(params, batch, f, t0, t1, d_dim, n_steps=16):
    """Drop-in for ofdft_normflows.jax_ode.neural_ode_score (jax_ode.py:51-110); ``f`` is the reference's own Flax module
    (``Gen_EqvFlow``: its static fields xyz_nuclei / z_one_hot / bool_neg are read; ``Gen_CNFSimpleMLP``: features / bool_neg)."""
    theta, _ = ravel_pytree(params)                 # key-sorted: last_layer.{bias,kernel}, layers_0.., layers_1..
    theta, batch = theta.astype(jnp.float32), batch.astype(jnp.float32)
    if _is_gnn(f):
        y1 = make_gnn_flow(f.xyz_nuclei, f.z_one_hot, bool_neg=f.bool_neg, jets=3, n_steps=n_steps, t0=t0, t1=t1)(theta, batch)
    else:
        y1 = make_mlp1d_flow(f.features, bool_neg=f.bool_neg, n_steps=n_steps, t0=t0, t1=t1)(theta, batch)
    return y1[:, :d_dim], y1[:, d_dim:d_dim + 1], y1[:, d_dim + 1:]


def neural_ode(params, batch, f, t0, t1, d_dim, n_steps=16):
    """Drop-in for ofdft_normflows.jax_ode.neural_ode (jax_ode.py:9-49)."""
    theta, _ = ravel_pytree(params)
    theta, batch = theta.astype(jnp.float32), batch.astype(jnp.float32)
    if _is_gnn(f):
        y1 = make_gnn_flow(f.xyz_nuclei, f.z_one_hot, bool_neg=f.bool_neg, jets=2, n_steps=n_steps, t0=t0, t1=t1)(theta, batch)
        return y1[:, :d_dim], y1[:, d_dim:]
    # 1-D: the adjoint is instantiated for full rows; a zero score column rides along (its cotangent stays zero)
    padded = jnp.concatenate([batch, jnp.zeros((batch.shape[0], 1), jnp.float32)], 1)
    y1 = make_mlp1d_flow(f.features, bool_neg=f.bool_neg, n_steps=n_steps, t0=t0, t1=t1)(theta, padded)
    return y1[:, :1], y1[:, 1:2]


# ---- functionals ------------------------------------------------------------------------------------------------------------
def make_energy_terms(*, Ne, d=3, kin="tf-w", hart="hartree", nuc="nuclei_potential", x="lda", c="", coords=None, z=None,
                      R=0.0, Z_alpha=0.0, Z_beta=0.0):
    """``terms(y1) -> float64[8]`` = per-term MEANS [kin, vnuc, hart, x, c, total, 0, 0] of the t1 state rows (2*bs, 2d+1)
    (the functional calls + jnp.mean of `loss`, H20_mol_ofdft_min_equiv_flows.py:150-173), with a custom VJP: the fused kernel
    produces the values and d(total)/d y1 in one pass, so the VJP is ``g[5] * ybar1`` (plus nothing for the per-term slots --
    the drivers differentiate the total only; a cotangent on another slot raises)."""
    coords_a = tuple(float(v) for v in np.asarray(coords if coords is not None else [], np.float32).reshape(-1))
    z_a = tuple(float(v) for v in np.asarray(z if z is not None else [], np.float32).reshape(-1))
    attrs = dict(d=np.int32(d), kin=np.int32(KIN[kin]), hart=np.int32(HART[hart]), nuc=np.int32(NUC[nuc]), x=np.int32(XC_X[x]),
                 c=np.int32(XC_C[c]), Ne=np.float32(Ne), R=np.float32(R), Z_alpha=np.float32(Z_alpha), Z_beta=np.float32(Z_beta),
                 coords=np.asarray(coords_a, np.float32), z=np.asarray(z_a, np.float32))

    def call(y1, want_bar):
        bs = y1.shape[0] // 2
        out = (jax.ShapeDtypeStruct((8,), jnp.float64),
               jax.ShapeDtypeStruct(y1.shape if want_bar else (1,), jnp.float32),
               _u8(_abi.ofdft_functionals_scratch_bytes(_i64(bs))))
        sums, ybar1, _ = jax.ffi.ffi_call("OfdftFunctionals", out)(y1, **attrs)
        return sums, ybar1

    @jax.custom_vjp
    def terms(y1):
        return call(y1, False)[0]

    def fwd(y1):
        sums, ybar1 = call(y1, True)
        return sums, ybar1

    def bwd(ybar1, g):
        return (g[5].astype(jnp.float32) * ybar1,)

    terms.defvjp(fwd, bwd)
    return terms


def make_hartree_allpairs(*, Ne, kind="hartree"):
    """``pairs(x (n,d), xp (m,d)) -> mean over all n*m cross pairs`` (north-star generalisation of functionals.py:463-486 /
    439-461), custom VJP from the kernel's own forces."""
    d = 3 if kind == "hartree" else 1
    attrs = dict(kind=np.int32(HART[kind]), d=np.int32(d), Ne=np.float32(Ne), norm_pairs=np.float64(0.0))

    def call(x, xp, want_grad):
        n, m = x.shape[0], xp.shape[0]
        sb = _abi.ofdft_hartree_allpairs_scratch_bytes(_i64(n), _i64(m))
        out = (jax.ShapeDtypeStruct((1,), jnp.float64),
               jax.ShapeDtypeStruct((n, d) if want_grad else (1, 1), jnp.float32),
               jax.ShapeDtypeStruct((m, d) if want_grad else (1, 1), jnp.float32), _u8(sb))
        e, xb, xpb, _ = jax.ffi.ffi_call("OfdftHartreeAllpairs", out)(x, xp, **attrs)
        return e[0], xb, xpb

    @jax.custom_vjp
    def pairs(x, xp):
        return call(x, xp, False)[0]

    def fwd(x, xp):
        e, xb, xpb = call(x, xp, True)
        return e, (xb, xpb)

    def bwd(res, g):
        xb, xpb = res
        g = g.astype(jnp.float32)
        return g * xb, g * xpb

    pairs.defvjp(fwd, bwd)
    return pairs
