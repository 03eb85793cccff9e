"""CPU execution of the general-depth GNN kernels (csrc/gnn_general.cu compiled unchanged for the host by
tests/hostsim/gnn_general_hostsim.cpp) against the float64 oracle: forward integrator, single RHS evaluation and the
discrete adjoint for 1 and 3 hidden layers (OFDFT_NF.py:320-326), mixed jets per segment and a ragged tail tile."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import ofdft_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "gnn_general_hostsim.cpp")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    if not os.path.isfile(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    out = str(tmp_path_factory.mktemp("hostsim") / "hostsim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", f"-I{CUDA_INC}", "-o", out, SRC], check=True)
    return C.CDLL(out)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _setup(L, mol_name, B, seed):
    rng = np.random.default_rng(seed)
    mol = O.molecule(mol_name)
    nt = mol["onehot"].shape[1]
    p = O.init_params(rng, 2 + nt, (64,) * L, 1, last_scale=0.5)
    theta = O.flatten_params(p).astype(np.float32)
    fld = O.GNNField(O.unflatten_params(theta.astype(np.float64), 2 + nt, (64,) * L), mol["coords"], mol["onehot"], bool_neg=True)
    batch = O.batch_generator_3d(rng, B // 2, mol["coords"], mol["z"]).astype(np.float32)[:B]
    return mol, nt, theta, fld, batch


@pytest.mark.parametrize("L", [1, 3])
def test_general_depth_forward_and_adjoint(sim, L):
    B, n_steps = 150, 2                       # two tiles, the second one ragged (22 rows)
    mol, nt, theta, fld, batch = _setup(L, "H2O", B, 10 + L)
    assert sim.hostsim_theta_size(nt, L) == theta.size
    Na = mol["coords"].shape[0]
    nuc = np.ascontiguousarray(mol["coords"], dtype=np.float32)
    oh = np.ascontiguousarray(mol["onehot"], dtype=np.float32)
    y1 = np.zeros_like(batch)
    ckpt = np.zeros(n_steps * 4 * 6 * B, dtype=np.float32)
    f32 = C.c_float
    rc = sim.hostsim_gnn_general_fwd(_p(theta), _p(nuc), _p(oh), _p(batch), _p(y1), _p(ckpt), C.c_longlong(B), C.c_longlong(B), 3, 3,
                                     7, Na, nt, L, n_steps, f32(0.0), f32(1.0), f32(-1.0), 0, f32(0.0))
    assert rc == 0
    fn = lambda y, t: fld.rhs(t, y, True)
    y_ref, stages = O.odeint_rk4(fn, batch.astype(np.float64), 0.0, 1.0, n_steps, keep=True)
    scale = np.abs(y_ref).max(0)
    assert (np.abs(y1 - y_ref) / scale).max() < 2e-5
    # one evaluation of the right-hand side at t = 0.3
    dy = np.zeros_like(batch)
    sim.hostsim_gnn_general_fwd(_p(theta), _p(nuc), _p(oh), _p(batch), _p(dy), None, C.c_longlong(B), C.c_longlong(B), 3, 3,
                                7, Na, nt, L, 1, f32(0.0), f32(1.0), f32(-1.0), 1, f32(0.3))
    dy_ref = fld.rhs(0.3, batch.astype(np.float64), True)
    assert (np.abs(dy - dy_ref) / np.abs(dy_ref).max(0)).max() < 2e-5
    # discrete adjoint with a random cotangent
    rng = np.random.default_rng(5)
    ybar1 = rng.standard_normal((B, 7)).astype(np.float32)
    n_tiles = (B + 127) // 128
    gpart = np.zeros((n_tiles, theta.size), dtype=np.float32)
    ybar0 = np.zeros_like(batch)
    rc = sim.hostsim_gnn_general_bwd(_p(theta), _p(nuc), _p(oh), _p(ckpt), _p(ybar1), _p(ybar0), _p(gpart), C.c_longlong(B),
                                     C.c_longlong(B), 3, 3, 7, Na, nt, L, n_steps, f32(0.0), f32(1.0), f32(-1.0))
    assert rc == 0
    yb0_ref, tbar_ref = O.rk4_adjoint(fld, stages, 0.0, 1.0, ybar1.astype(np.float64), True)
    tbar = gpart.astype(np.float64).sum(0)
    assert np.linalg.norm(tbar - tbar_ref) / np.linalg.norm(tbar_ref) < 1e-4
    lay = O.unflatten_params(np.arange(theta.size, dtype=np.float64), 2 + nt, (64,) * L)   # per-tensor check
    for w, b in lay.hidden + [lay.last]:
        for idx in (w.astype(int).ravel(), b.astype(int).ravel()):
            ref = tbar_ref[idx]
            if np.linalg.norm(ref) > 0:
                assert np.linalg.norm(tbar[idx] - ref) / np.linalg.norm(ref) < 1e-4
    assert (np.abs(ybar0 - yb0_ref) / np.abs(yb0_ref).max(0)).max() < 1e-4


def test_general_depth_mixed_jets(sim):
    """rows [0, n_a) carry all jets, rows [n_a, B) only x (the training layout)."""
    L, B, n_a, n_steps = 3, 96, 40, 2
    mol, nt, theta, fld, batch = _setup(L, "H2", B, 3)
    Na = mol["coords"].shape[0]
    nuc = np.ascontiguousarray(mol["coords"], dtype=np.float32)
    oh = np.ascontiguousarray(mol["onehot"], dtype=np.float32)
    y1 = np.zeros_like(batch)
    ckpt = np.zeros(n_steps * 4 * 6 * B, dtype=np.float32)
    f32 = C.c_float
    sim.hostsim_gnn_general_fwd(_p(theta), _p(nuc), _p(oh), _p(batch), _p(y1), _p(ckpt), C.c_longlong(B), C.c_longlong(n_a), 3, 1,
                                7, Na, nt, L, n_steps, f32(0.0), f32(1.0), f32(-1.0), 0, f32(0.0))
    ya, st_a = O.odeint_rk4(lambda y, t: fld.rhs(t, y, True), batch[:n_a].astype(np.float64), 0.0, 1.0, n_steps, keep=True)
    yb, st_b = O.odeint_rk4(lambda y, t: fld.rhs(t, y, False)[:, :3], batch[n_a:, :3].astype(np.float64), 0.0, 1.0, n_steps, keep=True)
    assert np.abs(y1[:n_a] - ya).max() < 5e-5 and np.abs(y1[n_a:, :3] - yb).max() < 5e-5
    rng = np.random.default_rng(8)
    ybar1 = rng.standard_normal((B, 7)).astype(np.float32)
    gpart = np.zeros((2, theta.size), dtype=np.float32)
    sim.hostsim_gnn_general_bwd(_p(theta), _p(nuc), _p(oh), _p(ckpt), _p(ybar1), None, _p(gpart), C.c_longlong(B),
                                C.c_longlong(n_a), 3, 1, 7, Na, nt, L, n_steps, f32(0.0), f32(1.0), f32(-1.0))
    _, ta = O.rk4_adjoint(fld, st_a, 0.0, 1.0, ybar1[:n_a].astype(np.float64), True)

    class XOnly:
        def rhs_vjp(self, t, y, kbar, with_score):
            y4 = np.concatenate([y, np.zeros((y.shape[0], 1))], 1)
            k4 = np.concatenate([kbar, np.zeros((y.shape[0], 1))], 1)
            yb_, g = fld.rhs_vjp(t, y4, k4, False)
            return yb_[:, :3], g
    _, tb = O.rk4_adjoint(XOnly(), st_b, 0.0, 1.0, ybar1[n_a:, :3].astype(np.float64), False)
    ref = ta + tb
    tbar = gpart.astype(np.float64).sum(0)
    assert np.linalg.norm(tbar - ref) / np.linalg.norm(ref) < 1e-4
