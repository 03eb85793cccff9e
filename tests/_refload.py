"""Import the reference's UNMODIFIED hot-path source files from /root/reference under the test-only jax/flax/distrax
stand-ins of tests/_jaxstub (see tests/_jaxstub/README.md).  Test infrastructure only.

    ref = load_reference()            # None when the reference checkout is absent (e.g. on the GPU box)
    ref.functionals.lda(den, score, Ne); ref.jax_ode.neural_ode_score(...); ref.equiv_flows.Gen_EqvFlow(...)

The package `__init__` of the reference (ofdft_normflows/__init__.py:1-6) also imports dft_distrax (PySCF), which is
outside the hot path; a package shell with the reference directory as `__path__` is registered instead so that
`from ofdft_normflows.utils import *` (functionals.py:10) resolves without executing it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("OFDFT_REFERENCE_ROOT", "/root/reference")
STUB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_jaxstub")
MODULES = ("utils", "functionals", "jax_ode", "equiv_flows", "cn_flows", "promolecular_distrax", "hgh_pseudopotentials")

_cache = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "ofdft_normflows", "functionals.py"))


def load_reference():
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        return None
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "__version__", "").endswith("torchstub"):
        raise RuntimeError("a real jax is already imported; the stub loader is for images without JAX")
    if STUB_DIR not in sys.path:
        sys.path.insert(0, STUB_DIR)
    pkg = types.ModuleType("ofdft_normflows")
    pkg.__path__ = [os.path.join(REF_ROOT, "ofdft_normflows")]
    sys.modules.setdefault("ofdft_normflows", pkg)
    ns = types.SimpleNamespace(root=REF_ROOT)
    for m in MODULES:
        setattr(ns, m, importlib.import_module(f"ofdft_normflows.{m}"))
    import jax
    ns.jax = jax
    _cache = ns
    return ns
