"""ctypes binding of the C-ABI shared library (include/ofdft_b200.h).

The product path has no CPU fallback: if the CUDA library is missing or fails to load, importing any
compute entry point raises.  ``OFDFT_B200_LIB`` overrides the library path (used to A/B kernel builds).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_LIB = None

vp, i32, i64, f32, szt = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class EnergySpecC(C.Structure):
    """``ofdft_energy_spec`` (include/ofdft_b200.h)."""
    _fields_ = [("d", i32), ("kin", i32), ("hart", i32), ("nuc", i32), ("x", i32), ("c", i32), ("n_nuclei", i32),
                ("Ne", f32), ("R", f32), ("Z_alpha", f32), ("Z_beta", f32)]


# name -> (restype, argtypes); every symbol include/ofdft_b200.h declares
SIGNATURES = {
    "ofdft_b200_abi_version": (C.c_int, []),
    "ofdft_b200_check_device": (C.c_int, []),
    "ofdft_b200_strerror": (C.c_char_p, [C.c_int]),
    "ofdft_cnf_gnn_ckpt_bytes": (szt, [i64, i32]),
    "ofdft_cnf_gnn_fwd_scratch_bytes": (szt, []),
    "ofdft_cnf_gnn_act_ckpt_bytes": (szt, [i64, i64, i32, i32, i32, i32]),
    "ofdft_cnf_gnn_theta_size": (i64, [i32, i32]),
    "ofdft_cnf_gnn_fwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, szt, i64, i64, i32, i32, i32, i32, i32, i32, i32, i32,
                                    f32, f32, f32, i32]),
    "ofdft_cnf_gnn_bwd_scratch_bytes": (szt, [i64, i64, i32, i32]),
    "ofdft_cnf_gnn_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, szt, i64, i64, i32, i32, i32, i32, i32, i32, i32,
                                    f32, f32, f32, i32]),
    "ofdft_cnf_gnn_rhs": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, szt, i64, i32, i32, i32, i32, i32, f32, f32, i32]),
    "ofdft_cnf_mlp1d_ckpt_bytes": (szt, [i64, i32]),
    "ofdft_cnf_mlp1d_act_ckpt_bytes": (szt, [i64, i32, i32, i32]),
    "ofdft_cnf_mlp1d_scratch_bytes": (szt, [i64, i32, i32, i32]),
    "ofdft_cnf_mlp1d_fwd": (C.c_int, [vp, vp, vp, vp, vp, vp, szt, i64, i32, i32, i32, i32, i32, f32, f32, f32, i32]),
    "ofdft_cnf_mlp1d_rhs": (C.c_int, [vp, vp, vp, vp, vp, szt, i64, i32, i32, i32, i32, f32, f32]),
    "ofdft_cnf_mlp1d_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, szt, i64, i32, i32, i32, i32, i32, f32, f32, f32, i32]),
    "ofdft_functionals_scratch_bytes": (szt, [i64]),
    "ofdft_functionals_fwd_bwd": (C.c_int, [vp, C.POINTER(EnergySpecC), vp, vp, vp, i32, i64, vp, vp, vp, szt]),
    "ofdft_functionals_integrands": (C.c_int, [vp, C.POINTER(EnergySpecC), vp, vp, vp, i32, i64, vp]),
    "ofdft_functional_eval": (C.c_int, [vp, i32, C.POINTER(EnergySpecC), vp, vp, vp, vp, i64, vp, vp, vp]),
    "ofdft_hartree_allpairs_scratch_bytes": (szt, [i64, i64]),
    "ofdft_hartree_allpairs": (C.c_int, [vp, i32, i32, f32, vp, i32, vp, i32, i64, i64, C.c_double, vp, vp, i32, vp, i32, i32, vp, szt]),
    "ofdft_promolecular_logp_score": (C.c_int, [vp, vp, vp, i32, vp, i32, i64, vp, i32, i32, vp, vp]),
    "ofdft_batch_generator_3d": (C.c_int, [vp, C.c_uint64, C.c_uint64, vp, vp, i32, i64, vp]),
    "ofdft_batch_generator_1d": (C.c_int, [vp, C.c_uint64, C.c_uint64, i64, vp]),
    "ofdft_rmsprop_clip_step": (C.c_int, [vp, vp, vp, vp, i64, f32, f32, f32, f32]),
    "ofdft_adam_clip_step": (C.c_int, [vp, vp, vp, vp, vp, i64, i64, f32, f32, f32, f32, f32]),
    "ofdft_ema_update": (C.c_int, [vp, vp, vp, vp, i32, i64, C.c_double]),
    "ofdft_microbench_ffma": (C.c_int, [vp, vp, i32, C.POINTER(C.c_double)]),
    "ofdft_microbench_sfu": (C.c_int, [vp, vp, i32, C.POINTER(C.c_double)]),
    "ofdft_tc_gemm_test": (C.c_int, [vp, vp, vp, vp, i32, i32]),
    "ofdft_tc_mma_rate": (C.c_int, [vp, vp, i32, i32, i32, i32, i32]),
}


def library_path() -> Path:
    env = os.environ.get("OFDFT_B200_LIB")
    return Path(env) if env else _HERE / "libofdft_b200.so"


def load():
    """Load (once) and return the ctypes handle; raises RuntimeError if the CUDA library is absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists():
        raise RuntimeError(
            f"ofdft_nflows_b200: CUDA library {path} not found -- build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the build is stale: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.ofdft_b200_abi_version() != 3:
        raise RuntimeError("ofdft_nflows_b200: ABI version mismatch between _lib.py and libofdft_b200.so")
    _LIB = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ofdft_b200_strerror(rc).decode()
        raise RuntimeError(f"{what or 'ofdft_b200 call'} failed: {msg} (code {rc})")
