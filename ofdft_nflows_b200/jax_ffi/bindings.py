"""jax.ffi + jax.custom_vjp drop-ins for ofdft_normflows (import-guarded: JAX is NOT installed in the build image,
so this module is exercised only where JAX >= 0.4.38 and libofdft_b200_xla.so (xla_ffi_shim.cc) are available).

    from ofdft_nflows_b200.jax_ffi.bindings import neural_ode_score        # same signature as jax_ode.py:51
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
for _name in ("OfdftGnnFwd", "OfdftGnnBwd", "OfdftFunctionals"):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_xla, _name)), platform="CUDA")
_abi.ofdft_cnf_gnn_ckpt_bytes.restype = ctypes.c_size_t
_abi.ofdft_cnf_gnn_act_ckpt_bytes.restype = ctypes.c_size_t
_abi.ofdft_cnf_gnn_bwd_scratch_bytes.restype = ctypes.c_size_t


def _gnn_fwd_call(theta, nuclei, onehot, y0, *, n_a, jets_a, jets_b, n_steps, t0, t1, y0sign, act_ckpt=True):
    B, w = y0.shape
    Na = nuclei.shape[0]
    ckpt_n = _abi.ofdft_cnf_gnn_ckpt_bytes(ctypes.c_int64(B), n_steps) // 4
    act_n = (_abi.ofdft_cnf_gnn_act_ckpt_bytes(ctypes.c_int64(B), ctypes.c_int64(n_a), jets_a, jets_b, Na, n_steps) // 4
             if act_ckpt else 1)
    out = (jax.ShapeDtypeStruct((B, w), jnp.float32), jax.ShapeDtypeStruct((ckpt_n,), jnp.float32),
           jax.ShapeDtypeStruct((act_n,), jnp.float32), jax.ShapeDtypeStruct((256,), jnp.uint8))
    y1, ckpt, act, _ = jax.ffi.ffi_call("OfdftGnnFwd", out)(
        theta, nuclei, onehot, y0, n_a=np.int64(n_a), jets_a=np.int32(jets_a), jets_b=np.int32(jets_b),
        n_steps=np.int32(n_steps), t0=np.float32(t0), t1=np.float32(t1), y0sign=np.float32(y0sign))
    return y1, ckpt, act


def make_gnn_flow(nuclei, onehot, *, bool_neg, jets, n_steps=16, t0=0.0, t1=1.0):
    """Returns ``flow(theta_flat, y0) -> y1`` with a custom VJP (discrete adjoint kernel)."""
    y0sign = -1.0 if bool_neg else 1.0
    nuclei = jnp.asarray(nuclei, jnp.float32)
    onehot = jnp.asarray(onehot, jnp.float32)
    n_types = onehot.shape[1]

    @jax.custom_vjp
    def flow(theta, y0):
        return _gnn_fwd_call(theta, nuclei, onehot, y0, n_a=y0.shape[0], jets_a=jets, jets_b=jets, n_steps=n_steps,
                             t0=t0, t1=t1, y0sign=y0sign)[0]

    def fwd(theta, y0):
        y1, ckpt, act = _gnn_fwd_call(theta, nuclei, onehot, y0, n_a=y0.shape[0], jets_a=jets, jets_b=jets,
                                      n_steps=n_steps, t0=t0, t1=t1, y0sign=y0sign)
        return y1, (theta, ckpt, act)

    def bwd(res, ybar1):
        theta, ckpt, act = res
        B, w = ybar1.shape
        sb = _abi.ofdft_cnf_gnn_bwd_scratch_bytes(ctypes.c_int64(B), ctypes.c_int64(B), n_types)
        out = (jax.ShapeDtypeStruct(theta.shape, jnp.float32), jax.ShapeDtypeStruct((B, w), jnp.float32),
               jax.ShapeDtypeStruct((sb,), jnp.uint8))
        tbar, ybar0, _ = jax.ffi.ffi_call("OfdftGnnBwd", out)(
            theta, nuclei, onehot, ckpt, act, ybar1, n_a=np.int64(B), jets_a=np.int32(jets), jets_b=np.int32(jets),
            n_steps=np.int32(n_steps), t0=np.float32(t0), t1=np.float32(t1), y0sign=np.float32(y0sign))
        return tbar, ybar0

    flow.defvjp(fwd, bwd)
    return flow


def neural_ode_score(params, batch, f, t0, t1, d_dim, n_steps=16):
    """Drop-in for ofdft_normflows.jax_ode.neural_ode_score (jax_ode.py:51-110); ``f`` is the reference's own Flax
    module ``Gen_EqvFlow`` (only its static fields xyz_nuclei / z_one_hot / bool_neg are read)."""
    theta, _ = ravel_pytree(params)                 # key-sorted: last_layer.{bias,kernel}, layers_0.., layers_1..
    flow = make_gnn_flow(f.xyz_nuclei, f.z_one_hot, bool_neg=f.bool_neg, jets=3, n_steps=n_steps, t0=t0, t1=t1)
    y1 = flow(theta.astype(jnp.float32), batch.astype(jnp.float32))
    return y1[:, :d_dim], y1[:, d_dim:d_dim + 1], y1[:, d_dim + 1:]


def neural_ode(params, batch, f, t0, t1, d_dim, n_steps=16):
    """Drop-in for ofdft_normflows.jax_ode.neural_ode (jax_ode.py:9-49)."""
    theta, _ = ravel_pytree(params)
    flow = make_gnn_flow(f.xyz_nuclei, f.z_one_hot, bool_neg=f.bool_neg, jets=2, n_steps=n_steps, t0=t0, t1=t1)
    y1 = flow(theta.astype(jnp.float32), batch.astype(jnp.float32))
    return y1[:, :d_dim], y1[:, d_dim:]
