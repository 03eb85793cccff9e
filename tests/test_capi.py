"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/ofdft_b200.h declares, and the
size helpers / argument validation behave (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re
import shutil
import subprocess
from pathlib import Path

import pytest

from ofdft_nflows_b200 import _lib

ROOT = Path(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HEADER = os.path.join(ROOT, "include", "ofdft_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ofdft_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ofdft_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_strerror():
    lib = _lib.load()
    assert lib.ofdft_b200_abi_version() == 3
    assert b"invalid" in lib.ofdft_b200_strerror(-1)
    assert b"unsupported" in lib.ofdft_b200_strerror(-2)
    assert b"scratch" in lib.ofdft_b200_strerror(-3)
    assert lib.ofdft_b200_strerror(0) == b"ok"


def test_size_helpers():
    lib = _lib.load()
    assert lib.ofdft_cnf_gnn_ckpt_bytes(131072, 16) == 16 * 4 * 6 * 131072 * 4
    assert lib.ofdft_cnf_gnn_fwd_scratch_bytes() >= 4
    n_theta = 4289 + 64 * 5
    assert lib.ofdft_cnf_gnn_theta_size(3, 2) == n_theta
    assert lib.ofdft_cnf_gnn_theta_size(3, 1) == n_theta - 64 - 4096 and lib.ofdft_cnf_gnn_theta_size(3, 3) == n_theta + 64 + 4096
    assert lib.ofdft_cnf_gnn_theta_size(3, 4) == 0
    assert lib.ofdft_cnf_gnn_bwd_scratch_bytes(131072, 65536, 3, 2) >= 1024 * n_theta * 4
    assert lib.ofdft_cnf_gnn_bwd_scratch_bytes(131072, 65536, 3, 3) >= 1024 * (n_theta + 4160) * 4
    assert lib.ofdft_cnf_gnn_act_ckpt_bytes(131072, 65536, 3, 1, 3, 16) == 16 * 4 * 3 * 64 * (3 * 65536 + 65536) * 4
    assert lib.ofdft_cnf_gnn_act_ckpt_bytes(300, 150, 3, 1, 3, 2) == 2 * 4 * 3 * 64 * (3 * 256 + 256) * 4   # rows padded to 128
    assert lib.ofdft_functionals_scratch_bytes(65536) >= (65536 // 256) * 5 * 8
    assert lib.ofdft_hartree_allpairs_scratch_bytes(1000, 777) > 0


def test_path_selectors_of_the_python_mirror_match_the_header():
    """ops.PATH_* / JETS_* are the OFDFT_PATH_* / OFDFT_JETS_* values of include/ofdft_b200.h (the kernel path is an argument of
    the ABI, not an environment variable); the bwd scratch holds the prepared weight operands of the depth-generic kernels."""
    from ofdft_nflows_b200 import ops
    src = open(HEADER).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(OFDFT_(?:PATH|JETS)_\w+)\s+(\d+)", src)}
    assert defs["OFDFT_PATH_DEFAULT"] == ops.PATH_DEFAULT and defs["OFDFT_PATH_FFMA"] == ops.PATH_FFMA
    assert defs["OFDFT_PATH_FFMA_DW"] == ops.PATH_FFMA_DW and defs["OFDFT_PATH_ACT_CKPT"] == ops.PATH_ACT_CKPT
    assert defs["OFDFT_PATH_DEEP"] == ops.PATH_DEEP == 8
    assert (defs["OFDFT_JETS_X"], defs["OFDFT_JETS_XLOGP"], defs["OFDFT_JETS_FULL"]) == (ops.JETS_X, ops.JETS_XLOGP, ops.JETS_FULL)
    lib = _lib.load()
    n3 = lib.ofdft_cnf_gnn_theta_size(3, 3)
    assert lib.ofdft_cnf_gnn_bwd_scratch_bytes(256, 128, 3, 3) >= 2 * n3 * 4 + 2 * 4 * 64 * 64 * 4     # 2 tiles + [layer][4][64 x 64]


def test_argument_validation_without_device():
    lib = _lib.load()
    # null pointers / bad sizes are rejected before any CUDA call
    N = None
    assert lib.ofdft_cnf_gnn_fwd(N, N, N, N, N, N, N, N, N, 0, 128, 128, 3, 3, 7, 7, 3, 3, 2, 16, 0.0, 1.0, -1.0, 0) == -1
    assert lib.ofdft_cnf_gnn_fwd(N, N, N, N, N, N, N, N, N, 0, 0, 0, 3, 3, 7, 7, 3, 3, 2, 16, 0.0, 1.0, -1.0, 0) == 0    # empty shard: no-op
    assert lib.ofdft_cnf_gnn_fwd(N, N, N, N, N, N, N, N, N, 0, -1, 0, 3, 3, 7, 7, 3, 3, 2, 16, 0.0, 1.0, -1.0, 0) == -1  # B < 0
    assert lib.ofdft_cnf_gnn_fwd(N, N, N, N, N, N, N, N, N, 0, 128, 128, 3, 3, 7, 7, 129, 3, 2, 16, 0.0, 1.0, -1.0, 0) == -2  # Na
    assert lib.ofdft_cnf_gnn_fwd(N, N, N, N, N, N, N, N, N, 0, 128, 128, 3, 3, 7, 7, 3, 3, 4, 16, 0.0, 1.0, -1.0, 0) == -2  # layers
    assert lib.ofdft_cnf_gnn_fwd(N, N, N, N, N, N, N, N, N, 0, 128, 128, 3, 3, 4, 7, 3, 3, 2, 16, 0.0, 1.0, -1.0, 0) == -1  # stride
    assert lib.ofdft_cnf_gnn_bwd(N, N, N, N, N, N, N, N, N, N, 0, 128, 128, 4, 3, 7, 3, 3, 2, 16, 0.0, 1.0, -1.0, 0) == -1  # jets
    assert lib.ofdft_cnf_mlp1d_fwd(N, N, N, N, N, N, 0, 0, 3, 3, 512, 3, 16, 0.0, 1.0, -1.0, 0) == 0          # empty shard: no-op
    assert lib.ofdft_cnf_mlp1d_rhs(N, N, N, N, N, 0, 16, 3, 3, 100, 3, 0.5, -1.0) == -2                         # width
    assert lib.ofdft_adam_clip_step(N, N, N, N, N, 10, 0, 1e-3, 0.9, 0.999, 1e-8, 1.0) == -1
    spec = _lib.EnergySpecC(2, 3, 1, 1, 1, 0, 0, 2.0, 0.0, 0.0, 0.0)   # d = 2 is not a thing
    assert lib.ofdft_functionals_fwd_bwd(None, C.byref(spec), None, None, None, 7, 16, None, None, None, 0) == -1
    assert lib.ofdft_hartree_allpairs(None, 1, 3, 2.0, None, 3, None, 3, 10, 10, 0.0, None, None, 3, None, 3, 0, None, 0) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("OFDFT_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_LIB", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


# ---- jax.ffi binding (xla_ffi_shim.cc / bindings.py): JAX is not installed here, so these are static / compile checks ----
SHIM = ROOT / "ofdft_nflows_b200" / "jax_ffi" / "xla_ffi_shim.cc"
BINDINGS = ROOT / "ofdft_nflows_b200" / "jax_ffi" / "bindings.py"
STUB_INC = ROOT / "tests" / "xla_ffi_stub"
CUDA_INC = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda")) / "include"


def _gxx(args, **kw):
    return subprocess.run(["g++", "-std=c++17", "-fsyntax-only", f"-I{STUB_INC}", f"-I{ROOT / 'include'}", f"-I{CUDA_INC}"] + args,
                          capture_output=True, text=True, **kw)


@pytest.mark.skipif(shutil.which("g++") is None or not (CUDA_INC / "cuda_runtime_api.h").exists(), reason="needs g++ and CUDA headers")
def test_xla_ffi_shim_compiles_against_the_c_abi(tmp_path):
    """Every XLA FFI handler calls its C-ABI entry point with the argument list include/ofdft_b200.h declares and is callable
    with the argument list its own binding declares (checked by the stand-in ffi.h of tests/xla_ffi_stub)."""
    r = _gxx([str(SHIM)])
    assert r.returncode == 0, r.stderr[-3000:]
    # the stand-in really checks: a handler whose binding lacks one attribute must not compile
    bad = tmp_path / "bad.cc"
    bad.write_text('#include <cuda_runtime_api.h>\n#include "xla/ffi/api/ffi.h"\nnamespace ffi = xla::ffi;\n'
                   "static ffi::Error Impl(cudaStream_t, ffi::Buffer<ffi::F32>, ffi::ResultBuffer<ffi::F32>, int32_t, float) { return ffi::Error::Success(); }\n"
                   "XLA_FFI_DEFINE_HANDLER_SYMBOL(Bad, Impl, ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()"
                   '.Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Attr<int32_t>("n"));\n')
    assert _gxx([str(bad)]).returncode != 0


def test_jax_bindings_are_consistent_with_the_shim_and_the_library():
    import ast
    import re
    src = BINDINGS.read_text()
    tree = ast.parse(src)
    consts = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id in ("FFI_TARGETS", "SIZE_HELPERS"):
            consts[node.targets[0].id] = [e.value for e in node.value.elts]
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+)", SHIM.read_text()))
    assert set(consts["FFI_TARGETS"]) == defined
    called = set(re.findall(r'ffi_call\("(\w+)"', src))
    assert called == defined                                   # every handler is wrapped, nothing else is called
    lib = _lib.load()
    for name in consts["SIZE_HELPERS"] + ["ofdft_cnf_gnn_theta_size"]:
        assert hasattr(lib, name), name
    # the attribute names of every binding appear in the Python wrapper
    for attr in set(re.findall(r'\.Attr<[^>]+>+\("(\w+)"\)', SHIM.read_text())):
        assert re.search(rf"\b{attr}=", src), attr
