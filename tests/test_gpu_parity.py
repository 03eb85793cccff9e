"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the float64 oracle on the
same seeded inputs, against the committed golden vectors (reference semantics: adaptive Dopri5 at 1e-7), and --
at BASELINE.json's full sizes -- through size-independent properties.

Tolerances (BASELINE.json north_star): every energy term <= 1e-5 relative; gradients <= 1e-4, measured as
||g - g_ref|| / ||g_ref|| per parameter tensor."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ofdft_oracle as O
from ofdft_nflows_b200 import ops, utils
from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow
from ofdft_nflows_b200.estimator import EnergyEstimator
from ofdft_nflows_b200.jax_ode import neural_ode, neural_ode_score

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
E_TOL, G_TOL = 1e-5, 1e-4


def dev(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float32), device="cuda")


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def tensor_slices(nt):
    I = 2 + nt
    return {"b3": (0, 1), "w3": (1, 65), "b1": (65, 129), "W1": (129, 129 + 64 * I), "b2": (129 + 64 * I, 193 + 64 * I),
            "W2": (193 + 64 * I, 4289 + 64 * I)}


def problem(molname, bs, seed, last_scale=0.5, biases=True):
    rng = np.random.default_rng(seed)
    mol = O.molecule(molname)
    nt = mol["onehot"].shape[1]
    p = O.init_params(rng, 2 + nt, (64, 64), 1, last_scale=last_scale * min(1.0, 3.0 / len(mol["z"])))
    if biases:
        for w, b in p.hidden:
            b[:] = 0.1 * rng.standard_normal(b.shape)
        p.last[1][:] = 0.05
    theta = O.flatten_params(p).astype(np.float32).astype(np.float64)
    batch = O.batch_generator_3d(rng, bs, mol["coords"], mol["z"]).astype(np.float32).astype(np.float64)
    fld = O.GNNField(O.unflatten_params(theta, 2 + nt, (64, 64)), mol["coords"], mol["onehot"], bool_neg=True)
    return mol, nt, theta, batch, fld


@pytest.mark.parametrize("molname,bs", [("H2", 128), ("H2O", 150), ("LiH", 65), ("C6H6", 70)])
def test_rhs_matches_oracle(molname, bs):
    mol, nt, theta, batch, fld = problem(molname, bs, 1)
    for y0sign, neg in ((-1.0, True), (1.0, False)):
        f = O.GNNField(fld.params, mol["coords"], mol["onehot"], bool_neg=neg)
        for jets, w in ((3, 7), (2, 4), (1, 3)):
            ref = f.rhs(0.3, batch, jets == 3)
            scale = np.abs(ref[:, :w]).max()
            for path in (ops.PATH_DEFAULT, ops.PATH_FFMA):
                dy = ops.gnn_rhs(dev(theta), dev(mol["coords"]), dev(mol["onehot"]), dev(batch), jets, 0.3, y0sign, path=path).cpu().numpy()
                assert np.abs(dy[:, :w] - ref[:, :w]).max() < 2e-6 * max(1.0, scale), path


@pytest.mark.parametrize("molname,bs,n_steps", [("H2", 128, 4), ("H2O", 150, 8), ("C6H6", 70, 4)])
def test_forward_matches_oracle_same_scheme(molname, bs, n_steps):
    """fp32 kernels vs the float64 oracle with the SAME fixed-step scheme (isolates arithmetic error)."""
    mol, nt, theta, batch, fld = problem(molname, bs, 2)
    B = 2 * bs
    ref = O.odeint_rk4(lambda y, t: fld.rhs(t, y, True), batch, 0.0, 1.0, n_steps)
    th, nuc, oh, y = dev(theta), dev(mol["coords"]), dev(mol["onehot"]), dev(batch)
    y1 = ops.gnn_fwd(th, nuc, oh, y, B, 3, 3, n_steps, 0.0, 1.0, -1.0)[0].cpu().numpy()
    # fp32 / 3xTF32 arithmetic: a few ulp of the largest state entry
    assert np.abs(y1[:, :4] - ref[:, :4]).max() < 2e-6 * max(1.0, np.abs(ref[:, :4]).max())
    assert np.abs(y1[:, 4:] - ref[:, 4:]).max() < 5e-6 * max(1.0, np.abs(ref[:, 4:]).max())
    # mixed launch: first half full jets, second half x only; ragged tile tails on both segments
    y1m = ops.gnn_fwd(th, nuc, oh, y, bs, 3, 1, n_steps, 0.0, 1.0, -1.0)[0].cpu().numpy()
    assert np.array_equal(y1m[:bs], y1[:bs])
    assert np.abs(y1m[bs:, :3] - ref[bs:, :3]).max() < 2e-6 * max(1.0, np.abs(ref[:, :3]).max())
    # neural_ode (no score channel), reverse direction t in [-1, 0] with bool_neg=False (rho_rev, H20_...py:132-137)
    frev = O.GNNField(fld.params, mol["coords"], mol["onehot"], bool_neg=False)
    z4 = np.concatenate([batch[:, :3], np.zeros((B, 1))], 1)
    ref_rev = O.odeint_rk4(lambda yy, t: frev.rhs(t, yy, False), z4, -1.0, 0.0, n_steps)
    y_rev = ops.gnn_fwd(th, nuc, oh, dev(z4), B, 2, 2, n_steps, -1.0, 0.0, 1.0)[0].cpu().numpy()
    assert np.abs(y_rev - ref_rev).max() < 2e-6 * max(1.0, np.abs(ref_rev).max())


@pytest.mark.parametrize("molname,bs,nuc", [("H2", 128, "nuclei_potential"), ("H2O", 150, "hgh"), ("C6H6", 70, "nuclei_potential")])
def test_training_step_matches_oracle_same_scheme(molname, bs, nuc):
    mol, nt, theta, batch, fld = problem(molname, bs, 3)
    n_steps = 4
    ospec = O.EnergySpec(Ne=mol["Ne"], kin="tf-w", hart="hartree", nuc=nuc, x="lda", coords=mol["coords"], z=mol["z"])
    e_ref, terms_ref, g_ref, _ = O.loss_and_grad(fld, ospec, batch, n_steps=n_steps)
    model = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=True)
    spec = ops.EnergySpec(Ne=mol["Ne"], d=3, kin="tf-w", hart="hartree", nuc=nuc, x="lda", coords=mol["coords"], z=mol["z"])
    # every kernel pair: tcgen05 forward + backward, tcgen05 forward + FP32 recompute backward, and the FP32-pipe
    # (PATH_FFMA) forward with its checkpoint-reading / recomputing backward kernels
    for skip, act, path in ((True, True, ops.PATH_DEFAULT), (True, False, ops.PATH_DEFAULT), (False, True, ops.PATH_DEFAULT),
                            (False, False, ops.PATH_DEFAULT), (True, True, ops.PATH_FFMA), (True, False, ops.PATH_FFMA)):
        est = EnergyEstimator(model, spec, n_steps=n_steps, skip_unused_jets=skip, act_ckpt=act, path=path)
        sums, tbar, _ = est.step_flat(dev(theta), dev(batch))
        sums = sums.cpu().numpy(); g = tbar.cpu().numpy()
        for i in range(4):
            assert abs(sums[i] - terms_ref[i]) <= E_TOL * abs(terms_ref[i]), (i, sums[i], terms_ref[i])
        assert abs(sums[5] - e_ref) <= E_TOL * abs(e_ref)
        for name, (a, b) in tensor_slices(nt).items():
            assert rel(g[a:b], g_ref[a:b]) <= G_TOL, (name, rel(g[a:b], g_ref[a:b]))


@pytest.mark.parametrize("path", sorted(p for p in glob.glob(os.path.join(GOLD, "*.npz"))
                                         if "1d" not in os.path.basename(p) and not os.path.basename(p).startswith("ref_")))
def test_golden_vectors_reference_semantics(path):
    """Committed fixtures = the reference's semantics (adaptive Dopri5 1e-7, float64, autodiff gradient).
    The fixed-step fp32 kernels must land within the north-star tolerances."""
    g = np.load(path)
    nt = g["onehot"].shape[1]
    model = Gen_EqvFlow(3, (64, 64), g["coords"], g["onehot"], bool_neg=True)
    spec = ops.EnergySpec(Ne=float(g["Ne"]), d=3, kin="tf-w", hart="hartree", nuc=str(g["nuc"]), x="lda", coords=g["coords"], z=g["z"])
    est = EnergyEstimator(model, spec, n_steps=32)
    sums, tbar, y1 = est.step_flat(dev(g["theta"]), dev(g["batch"]))
    sums = sums.cpu().numpy(); grad = tbar.cpu().numpy()
    for i in range(4):
        assert abs(sums[i] - g["terms"][i]) <= E_TOL * abs(g["terms"][i]), (i, sums[i], g["terms"][i])
    assert abs(sums[5] - float(g["energy"])) <= E_TOL * abs(float(g["energy"]))
    gref = g["grad"]
    for name, (a, b) in tensor_slices(nt).items():
        if np.linalg.norm(gref[a:b]) == 0.0:
            assert np.linalg.norm(grad[a:b]) == 0.0, name          # zero-init: only the last layer has a gradient
        else:
            assert rel(grad[a:b], gref[a:b]) <= G_TOL, (name, rel(grad[a:b], gref[a:b]))
    bs = g["batch"].shape[0] // 2
    assert np.abs(y1.cpu().numpy()[:bs, :3] - g["y1"][:bs, :3]).max() < 1e-4


def _synthetic_field(Na, nt, seed, feats=(64, 64)):
    """Random geometry with ``Na`` nuclei of ``nt`` species (not a reference molecule: exercises the kernel limits)."""
    rng = np.random.default_rng(seed)
    coords = 2.5 * rng.standard_normal((Na, 3))
    onehot = np.zeros((Na, nt)); onehot[np.arange(Na), rng.integers(0, nt, Na)] = 1.0
    p = O.init_params(rng, 2 + nt, feats, 1, last_scale=0.5 * min(1.0, 3.0 / Na))
    for w, b in p.hidden:
        b[:] = 0.1 * rng.standard_normal(b.shape)
    p.last[1][:] = 0.05
    theta = O.flatten_params(p).astype(np.float32).astype(np.float64)
    fld = O.GNNField(O.unflatten_params(theta, 2 + nt, feats), coords, onehot, bool_neg=True)
    return coords, onehot, theta, fld


@pytest.mark.parametrize("Na,nt,B,n_a,n_steps", [(1, 1, 1, 1, 1), (3, 2, 1, 0, 2), (16, 3, 130, 129, 2), (16, 4, 257, 0, 1),
                                                 (5, 0, 40, 40, 3), (74, 3, 70, 40, 1), (128, 8, 33, 20, 1)])
def test_edge_shapes_forward_and_adjoint_match_oracle(Na, nt, B, n_a, n_steps):
    """Single-row batches, an empty first / second segment, tile tails of one row, 74 nuclei of 3 species (the size of
    C27H46O, utils.py:327), the maximum nucleus and species counts (128 / 8), a species-free one-hot (n_types = 0) and a
    single RK4 step -- forward states and the discrete adjoint (both checkpoint modes) against the float64 oracle with
    the same scheme."""
    coords, onehot, theta, fld = _synthetic_field(Na, max(nt, 1), 11 + Na + B)
    if nt == 0:
        onehot = np.zeros((Na, 0))
        I = 2
        sl = tensor_slices(1)["W1"]
        keep = np.ones(theta.size, bool); keep[sl[0] + 64 * I: sl[1]] = False
        theta = theta[keep]
        fld = O.GNNField(O.unflatten_params(theta, I, (64, 64)), coords, onehot, bool_neg=True)
    rng = np.random.default_rng(5)
    y0 = np.concatenate([coords[rng.integers(0, Na, B)] + rng.standard_normal((B, 3)), rng.standard_normal((B, 4))], 1)
    y0 = y0.astype(np.float32).astype(np.float64)
    ybar1 = rng.standard_normal((B, 7))
    ybar1[n_a:, 3:] = 0.0                      # rows of the x-only segment have no logp / score cotangent
    fn = lambda y, t: fld.rhs(t, y, True)
    ref, stages = O.odeint_rk4(fn, y0, 0.0, 1.0, n_steps, keep=True)
    ybar0_ref, g_ref = O.rk4_adjoint(fld, stages, 0.0, 1.0, ybar1, True)
    th, nuc, oh = dev(theta), dev(coords), torch.zeros((Na, nt), device="cuda") if nt == 0 else dev(onehot)
    for want_act in (True, False):
        y1, ck = ops.gnn_fwd(th, nuc, oh, dev(y0), n_a, 3, 1, n_steps, 0.0, 1.0, -1.0, want_ckpt=True, want_act=want_act)
        y1 = y1.cpu().numpy()
        assert np.abs(y1[:n_a] - ref[:n_a]).max(initial=0.0) < 5e-6 * max(1.0, np.abs(ref).max())
        assert np.abs(y1[n_a:, :3] - ref[n_a:, :3]).max(initial=0.0) < 5e-6 * max(1.0, np.abs(ref).max())
        g, yb0 = ops.gnn_bwd(th, nuc, oh, ck, dev(ybar1), n_a, 3, 1, n_steps, 0.0, 1.0, -1.0, want_ybar0=True)
        assert rel(g.cpu().numpy(), g_ref) <= G_TOL, rel(g.cpu().numpy(), g_ref)
        yb0 = yb0.cpu().numpy()
        assert rel(yb0[:n_a], ybar0_ref[:n_a]) <= G_TOL if n_a else True
        assert rel(yb0[n_a:, :3], ybar0_ref[n_a:, :3]) <= G_TOL if n_a < B else True


@pytest.mark.parametrize("n_layers,path", [(3, ops.PATH_DEFAULT), (3, ops.PATH_FFMA), (1, ops.PATH_DEFAULT), (2, ops.PATH_DEEP)])
@pytest.mark.parametrize("Na,nt,B,n_a,n_steps,jets_a", [(1, 1, 1, 1, 1, 3), (3, 2, 130, 129, 2, 3), (16, 4, 257, 0, 1, 3), (74, 3, 70, 40, 1, 3),
                                                        (4, 2, 200, 77, 2, 2)])
def test_edge_shapes_other_depths_match_oracle(n_layers, path, Na, nt, B, n_a, n_steps, jets_a):
    """The edge shapes above for the --nn 1 / --nn 3 networks (OFDFT_NF.py:320-326) on both kernel families of depth 3 (tensor
    cores: gnn_deep_tc.cu; FP32: gnn_general.cu) and, as a cross-check, the depth-generic tensor-core kernels on (64, 64):
    forward states, one RHS evaluation and the discrete adjoint against the float64 oracle with the same scheme; one case
    carries (x, logp) only in its first segment (neural_ode, jax_ode.py:9-49)."""
    feats = (64,) * n_layers
    coords, onehot, theta, fld = _synthetic_field(Na, nt, 23 + Na + B, feats)
    rng = np.random.default_rng(6)
    y0 = np.concatenate([coords[rng.integers(0, Na, B)] + rng.standard_normal((B, 3)), rng.standard_normal((B, 4))], 1)
    y0 = y0.astype(np.float32).astype(np.float64)
    ybar1 = rng.standard_normal((B, 7))
    ybar1[n_a:, 3:] = 0.0                      # rows of the x-only segment have no logp / score cotangent
    if jets_a == 2:
        ybar1[:, 4:] = 0.0                     # no score channel anywhere
    score = jets_a == 3
    nc = 7 if score else 4
    fn = lambda y, t: fld.rhs(t, y, score)
    ref, stages = O.odeint_rk4(fn, y0[:, :nc], 0.0, 1.0, n_steps, keep=True)
    ybar0_ref, g_ref = O.rk4_adjoint(fld, stages, 0.0, 1.0, ybar1[:, :nc], score)
    th, nuc, oh = dev(theta), dev(coords), dev(onehot)
    y1, ck = ops.gnn_fwd(th, nuc, oh, dev(y0), n_a, jets_a, 1, n_steps, 0.0, 1.0, -1.0, want_ckpt=True, path=path)
    y1 = y1.cpu().numpy()
    assert np.abs(y1[:n_a, :nc] - ref[:n_a]).max(initial=0.0) < 5e-6 * max(1.0, np.abs(ref).max())
    assert np.abs(y1[n_a:, :3] - ref[n_a:, :3]).max(initial=0.0) < 5e-6 * max(1.0, np.abs(ref).max())
    g, yb0 = ops.gnn_bwd(th, nuc, oh, ck, dev(ybar1), n_a, jets_a, 1, n_steps, 0.0, 1.0, -1.0, want_ybar0=True, path=path)
    assert rel(g.cpu().numpy(), g_ref) <= G_TOL, rel(g.cpu().numpy(), g_ref)
    yb0 = yb0.cpu().numpy()
    assert rel(yb0[:n_a, :nc], ybar0_ref[:n_a]) <= G_TOL if n_a else True
    assert rel(yb0[n_a:, :3], ybar0_ref[n_a:, :3]) <= G_TOL if n_a < B else True
    dy = ops.gnn_rhs(th, nuc, oh, dev(y0), jets_a, 0.37, -1.0, path=path).cpu().numpy()
    dref = fld.rhs(0.37, y0[:, :nc], score)
    assert np.abs(dy[:, :nc] - dref).max() < 2e-5 * max(1.0, np.abs(dref).max())


def test_error_behaviour_of_the_c_abi():
    """Bad arguments are reported, never silently computed on: more nuclei than the kernels are instantiated for,
    a theta of the wrong length, host tensors."""
    coords, onehot, theta, _ = _synthetic_field(129, 2, 3)
    y = torch.zeros((4, 7), device="cuda")
    with pytest.raises(RuntimeError, match="code -2"):
        ops.gnn_fwd(dev(theta), dev(coords), dev(onehot), y, 4, 3, 3, 2, 0.0, 1.0, -1.0)
    with pytest.raises(NotImplementedError):
        Gen_EqvFlow(3, (64, 64), coords, onehot)
    with pytest.raises(NotImplementedError):
        Gen_EqvFlow(3, (64, 32), coords[:3], onehot[:3])
    coords, onehot, theta, _ = _synthetic_field(3, 2, 3)
    # an empty shard is not an error: no states, a zero gradient
    y1, ck = ops.gnn_fwd(dev(theta), dev(coords), dev(onehot), torch.zeros((0, 7), device="cuda"), 0, 3, 3, 2, 0.0, 1.0, -1.0,
                         want_ckpt=True, want_act=True)
    assert y1.shape == (0, 7)
    g, _ = ops.gnn_bwd(dev(theta), dev(coords), dev(onehot), ck, torch.zeros((0, 7), device="cuda"), 0, 3, 3, 2, 0.0, 1.0, -1.0)
    assert g.shape == (theta.size,) and float(g.abs().max()) == 0.0
    with pytest.raises(RuntimeError, match="code -1"):
        ops.gnn_fwd(dev(theta), dev(coords), dev(onehot), y, 5, 3, 3, 2, 0.0, 1.0, -1.0)        # n_a > B
    with pytest.raises(RuntimeError, match="code -1"):
        ops.gnn_fwd(dev(theta), dev(coords), dev(onehot), y, 4, 3, 3, 0, 0.0, 1.0, -1.0)        # n_steps = 0
    with pytest.raises(ValueError):
        ops.gnn_fwd(dev(theta)[:-1], dev(coords), dev(onehot), y, 4, 3, 3, 2, 0.0, 1.0, -1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gnn_fwd(torch.as_tensor(theta, dtype=torch.float32), dev(coords), dev(onehot), y, 4, 3, 3, 2, 0.0, 1.0, -1.0)


def test_chunked_backward_above_the_activation_checkpoint_cap():
    """When the layer-2 activation checkpoint of the whole batch exceeds the cap the estimator runs the step over row chunks
    of both halves (forward -> functionals -> adjoint per chunk, tensor-core backward kept, no extra forward pass): same
    energies, same gradient as the single-launch path."""
    mol, nt, theta, batch, fld = problem("H2O", 300, 7)
    model = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=True)
    spec = ops.EnergySpec(Ne=mol["Ne"], d=3, kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=mol["coords"], z=mol["z"])
    full = EnergyEstimator(model, spec, n_steps=4)
    per_row = ops.gnn_act_ckpt_bytes(128, 128, 3, 3, 3, 4) // 128
    small = EnergyEstimator(model, spec, n_steps=4, act_ckpt_max_bytes=130 * per_row)       # chunks of 128 rows
    s0, g0, _ = full.step_flat(dev(theta), dev(batch))
    s1, g1, _ = small.step_flat(dev(theta), dev(batch))
    assert np.allclose(s0.cpu().numpy(), s1.cpu().numpy(), rtol=1e-6)        # float32 cotangent scaling / chunk re-weighting
    assert rel(g1.cpu().numpy(), g0.cpu().numpy()) < 2e-6
    # all-pairs mode (every row is coupled to every partner): one launch, tensor-core recompute of the layer-2 jets
    fa = EnergyEstimator(model, spec, n_steps=4, hartree_mode="allpairs")
    sa = EnergyEstimator(model, spec, n_steps=4, hartree_mode="allpairs", act_ckpt_max_bytes=130 * per_row)
    s2, g2, _ = fa.step_flat(dev(theta), dev(batch))
    s3, g3, _ = sa.step_flat(dev(theta), dev(batch))
    assert torch.equal(s2, s3) and rel(g3.cpu().numpy(), g2.cpu().numpy()) < 5e-6
    ospec = O.EnergySpec(Ne=mol["Ne"], kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=mol["coords"], z=mol["z"])
    _, _, g_ref, _ = O.loss_and_grad(fld, ospec, batch, n_steps=4)
    for name, (a, b) in tensor_slices(nt).items():
        assert rel(g1.cpu().numpy()[a:b], g_ref[a:b]) <= G_TOL, name


def test_functionals_all_modes_match_oracle():
    rng = np.random.default_rng(4)
    mol = O.molecule("H2O")
    y1 = O.batch_generator_3d(rng, 300, mol["coords"], mol["z"]).astype(np.float32).astype(np.float64)
    for kin in ("tf", "w", "tf-w"):
        for nuc in ("nuclei_potential", "hgh"):
            for hart in ("hartree", "hartree_mt"):
                for x in ("lda", ""):
                    ospec = O.EnergySpec(Ne=10, kin=kin, hart=hart, nuc=nuc, x=x, coords=mol["coords"], z=mol["z"])
                    spec = ops.EnergySpec(Ne=10, d=3, kin=kin, hart=hart, nuc=nuc, x=x, coords=mol["coords"], z=mol["z"])
                    e_ref, terms_ref = O.energy_terms(ospec, y1, 3)
                    sums, yb = ops.functionals_fwd_bwd(spec, dev(y1))
                    sums = sums.cpu().numpy()
                    for i in range(4):
                        assert abs(sums[i] - terms_ref[i]) <= 1e-6 * max(abs(terms_ref[i]), 1e-30)
                    assert rel(yb.cpu().numpy(), O.energy_terms_vjp(ospec, y1, 3)) < 1e-5
                    integ = ops.functionals_integrands(spec, dev(y1)).cpu().numpy().astype(np.float64)
                    assert np.allclose(integ.mean(1)[:4], terms_ref[:4], rtol=1e-6, atol=1e-12)
    # 1-D functionals (LiH.py defaults)
    y1d = O.batch_generator_1d(rng, 257).astype(np.float32).astype(np.float64)
    ospec = O.EnergySpec(Ne=2, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
    spec = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
    e_ref, terms_ref = O.energy_terms(ospec, y1d, 1)
    sums, yb = ops.functionals_fwd_bwd(spec, dev(y1d))
    assert np.allclose(sums.cpu().numpy()[:3], terms_ref[:3], rtol=1e-6)
    assert rel(yb.cpu().numpy(), O.energy_terms_vjp(ospec, y1d, 1)) < 1e-5


def test_allpairs_hartree_training_step_matches_oracle():
    """hartree_mode='allpairs': the O(bs^2) cross-half estimator (north star) inside the fused step."""
    mol, nt, theta, batch, fld = problem("H2O", 150, 12)
    ospec = O.EnergySpec(Ne=10, kin="tf-w", hart="hartree_allpairs", nuc="nuclei_potential", x="lda", coords=mol["coords"], z=mol["z"])
    e_ref, terms_ref, g_ref, y1_ref = O.loss_and_grad(fld, ospec, batch, n_steps=4)
    model = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=True)
    spec = ops.EnergySpec(Ne=10, d=3, kin="tf-w", hart="hartree", nuc="nuclei_potential", x="lda", coords=mol["coords"], z=mol["z"])
    # the second half now needs no logp/score either: x-only jets stay valid
    est = EnergyEstimator(model, spec, n_steps=4, hartree_mode="allpairs")
    sums, tbar, _ = est.step_flat(dev(theta), dev(batch))
    sums = sums.cpu().numpy()
    for i in range(4):
        assert abs(sums[i] - terms_ref[i]) <= E_TOL * abs(terms_ref[i]), (i, sums[i], terms_ref[i])
    assert abs(sums[5] - e_ref) <= E_TOL * abs(e_ref)
    for name, (a, b) in tensor_slices(nt).items():
        assert rel(tbar.cpu().numpy()[a:b], g_ref[a:b]) <= G_TOL, name
    # all-pairs estimates the same double integral as the paired estimator, with lower variance
    e_pair = O.hartree_potential(y1_ref[:150, :3], y1_ref[150:, :3], 10).mean()
    assert abs(terms_ref[2] - e_pair) < 0.2 * abs(e_pair)


def test_xc_functionals_driver_defaults():
    """The reference drivers' default XC selection: --x dirac_b88_x_e --c vwn_c_e (H20_...py:241-244), pw92_c_e,
    and xc_1d (LiH.py:217) -- fused kernel terms, cotangents and a full training step."""
    rng = np.random.default_rng(9)
    mol = O.molecule("H2O")
    y1 = O.batch_generator_3d(rng, 200, mol["coords"], mol["z"]).astype(np.float32).astype(np.float64)
    for xname, cname in (("dirac_b88_x_e", "vwn_c_e"), ("b88_x_e", "pw92_c_e"), ("lda", "vwn_c_e")):
        ospec = O.EnergySpec(Ne=10, kin="tf-w", hart="hartree", nuc="nuclei_potential", x=xname, c=cname, coords=mol["coords"], z=mol["z"])
        spec = ops.EnergySpec(Ne=10, d=3, kin="tf-w", hart="hartree", nuc="nuclei_potential", x=xname, c=cname, coords=mol["coords"], z=mol["z"])
        e_ref, terms_ref = O.energy_terms(ospec, y1, 3)
        sums, yb = ops.functionals_fwd_bwd(spec, dev(y1))
        sums = sums.cpu().numpy()
        for i in range(5):
            assert abs(sums[i] - terms_ref[i]) <= 2e-6 * max(abs(terms_ref[i]), 1e-30), (xname, cname, i, sums[i], terms_ref[i])
        assert abs(sums[5] - e_ref) <= 2e-6 * abs(e_ref)
        assert rel(yb.cpu().numpy(), O.energy_terms_grad(ospec, y1, 3)) < 2e-5
    y1d = O.batch_generator_1d(rng, 300).astype(np.float32).astype(np.float64)
    ospec = O.EnergySpec(Ne=2, kin="tfw_1d", hart="softc", nuc="attraction", x="", c="xc_1d", R=0.7, Z_alpha=3, Z_beta=1)
    spec = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", c="xc_1d", R=0.7, Z_alpha=3, Z_beta=1)
    e_ref, terms_ref = O.energy_terms(ospec, y1d, 1)
    sums, yb = ops.functionals_fwd_bwd(spec, dev(y1d))
    assert np.allclose(sums.cpu().numpy()[:5], terms_ref, rtol=2e-6, atol=1e-12)
    assert rel(yb.cpu().numpy(), O.energy_terms_grad(ospec, y1d, 1)) < 2e-5
    # full training step with the H2O driver defaults (tf-w, nuclei_potential, hartree, dirac_b88_x_e, vwn_c_e)
    mol, nt, theta, batch, fld = problem("H2O", 130, 10)
    ospec = O.EnergySpec(Ne=10, kin="tf-w", hart="hartree", nuc="nuclei_potential", x="dirac_b88_x_e", c="vwn_c_e", coords=mol["coords"], z=mol["z"])
    spec = ops.EnergySpec(Ne=10, d=3, kin="tf-w", hart="hartree", nuc="nuclei_potential", x="dirac_b88_x_e", c="vwn_c_e", coords=mol["coords"], z=mol["z"])
    e_ref, terms_ref, g_ref, _ = O.loss_and_grad(fld, ospec, batch, n_steps=4)
    model = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=True)
    sums, tbar, _ = EnergyEstimator(model, spec, n_steps=4).step_flat(dev(theta), dev(batch))
    sums = sums.cpu().numpy()
    for i in range(5):
        assert abs(sums[i] - terms_ref[i]) <= E_TOL * abs(terms_ref[i]), (i, sums[i], terms_ref[i])
    for name, (a, b) in tensor_slices(nt).items():
        assert rel(tbar.cpu().numpy()[a:b], g_ref[a:b]) <= G_TOL, name
    # c-type factory signature (den, Ne)
    from ofdft_nflows_b200.functionals import _exchange_correlation
    den = np.exp(y1[:200, 3:4])
    ec = _exchange_correlation("vwn_c_e")(dev(den), 10).cpu().numpy()
    assert ec.shape == (200, 1) and np.allclose(ec, O.correlation_vwn_c_e(den.astype(np.float32).astype(np.float64), 10), rtol=5e-6)
    ex = _exchange_correlation("dirac_b88_x_e")(dev(den), dev(y1[:200, 4:7]), 10).cpu().numpy()
    d32 = den.astype(np.float32).astype(np.float64)
    assert np.allclose(ex, O.lda(d32, y1[:200, 4:7], 10) + O.b88_x_e(d32, y1[:200, 4:7], 10), rtol=5e-6)


def test_functional_factories_drop_in_and_autograd():
    """_kinetic/_hartree/_nuclear/_exchange_correlation(name)(...) with the reference signatures, (bs,1) outputs and
    gradients through torch.autograd: a `loss` written like H20_mol_ofdft_min_equiv_flows.py:150-173."""
    from ofdft_nflows_b200.functionals import _kinetic, _hartree, _nuclear, _exchange_correlation
    rng = np.random.default_rng(8)
    mol = O.molecule("H2O")
    bs = 257
    y = O.batch_generator_3d(rng, bs, mol["coords"], mol["z"]).astype(np.float32).astype(np.float64)
    den_np = np.exp(y[:bs, 3:4]); sc_np = y[:bs, 4:7]; x_np = y[:bs, :3]; xp_np = y[bs:, :3]
    den = dev(den_np).requires_grad_(True); score = dev(sc_np).requires_grad_(True)
    x = dev(x_np).requires_grad_(True); xp = dev(xp_np).requires_grad_(True)
    Ne = 10
    mol_info = {"coords": mol["coords"], "z": mol["z"]}
    e_t = _kinetic("tf-w")(den, score, Ne)
    e_h = _hartree("hartree")(x, xp, Ne)
    e_n = _nuclear("hgh")(x, Ne, mol_info)
    e_x = _exchange_correlation("lda")(den, score, Ne)
    assert e_t.shape == (bs, 1) and e_h.shape == (bs, 1) and e_n.shape == (bs, 1) and e_x.shape == (bs, 1)
    ref_t = O.thomas_fermi(den_np, sc_np, Ne) + O.weizsacker(den_np, sc_np, Ne)
    ref_h = O.hartree_potential(x_np, xp_np, Ne)
    ref_n = O.nuclei_potential_hgh(x_np, Ne, mol["coords"], mol["z"])
    ref_x = O.lda(den_np, sc_np, Ne)
    for got, ref in ((e_t, ref_t), (e_h, ref_h), (e_n, ref_n), (e_x, ref_x)):
        assert np.allclose(got.detach().cpu().numpy(), ref, rtol=2e-6, atol=1e-6 * np.abs(ref).max())
    energy = (e_t + e_n + e_h + e_x).mean()
    energy.backward()
    # cotangents == the fused kernel's (and the oracle's) d energy / d (x, logp, score)
    ospec = O.EnergySpec(Ne=Ne, kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=mol["coords"], z=mol["z"])
    g = O.energy_terms_vjp(ospec, y, 3)
    assert rel(x.grad.cpu().numpy(), g[:bs, :3]) < 1e-5 and rel(xp.grad.cpu().numpy(), g[bs:, :3]) < 1e-5
    assert rel(score.grad.cpu().numpy(), g[:bs, 4:7]) < 1e-5
    assert rel((den.grad * den).detach().cpu().numpy(), g[:bs, 3:4]) < 1e-5          # d/dlogp = den * d/dden
    # 1-D factories (LiH.py:118-138)
    y1 = O.batch_generator_1d(rng, 64).astype(np.float32).astype(np.float64)
    d1 = np.exp(y1[:64, 1:2]); s1 = y1[:64, 2:3]
    k1 = _kinetic("tfw_1d")(dev(d1), dev(s1), 2).cpu().numpy()
    assert np.allclose(k1, O.thomas_fermi_1D(d1, s1, 2) + O.weizsacker(d1, s1, 2), rtol=2e-6)
    h1 = _hartree("softc")(dev(y1[:64, :1]), dev(y1[64:, :1]), 2).cpu().numpy()
    assert np.allclose(h1, O.soft_coulomb(y1[:64, :1], y1[64:, :1], 2), rtol=2e-6)
    n1 = _nuclear("attraction")(dev(y1[:64, :1]), 0.7, 3, 1, 2).cpu().numpy()
    assert np.allclose(n1, O.attraction(y1[:64, :1], 0.7, 3, 1, 2), rtol=2e-6)
    with pytest.raises(NotImplementedError):
        _kinetic("kinetic")


def test_hartree_allpairs_matches_numpy():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1500, 7)); xp = rng.standard_normal((1111, 7)) + 0.3
    es, xb, xpb = ops.hartree_allpairs(dev(x), dev(xp), 10.0, "hartree")
    xf, xpf = x[:, :3].astype(np.float32).astype(np.float64), xp[:, :3].astype(np.float32).astype(np.float64)
    diff = xf[:, None, :] - xpf[None]
    zz = (diff ** 2).sum(-1) + 3e-5
    assert abs(es.item() - 50.0 * (1 / np.sqrt(zz)).mean()) <= 1e-6 * es.item()
    gref = (-50.0 * diff / zz[..., None] ** 1.5) / (1500 * 1111)
    assert rel(xb.cpu().numpy(), gref.sum(1)) < 1e-5 and rel(xpb.cpu().numpy(), -gref.sum(0)) < 1e-5
    # the paired estimator is the diagonal of the pair matrix: all-pairs on n == m reproduces its mean when fed row i only
    x1 = rng.standard_normal((700, 3)); x1p = rng.standard_normal((900, 3))
    es1, xb1, _ = ops.hartree_allpairs(dev(x1[:, :1].repeat(3, 1)), dev(x1p[:, :1].repeat(3, 1)), 2.0, "softc")
    a = x1[:, :1].astype(np.float32).astype(np.float64); b = x1p[:, :1].astype(np.float32).astype(np.float64)
    assert abs(es1.item() - 4.0 * (1 / np.sqrt(1 + (a - b.T) ** 2)).mean()) <= 1e-6 * es1.item()


def test_drop_in_neural_ode_score_autograd():
    """neural_ode_score(params, batch, f, t0, t1, d) with torch.autograd standing in for jax.custom_vjp."""
    mol, nt, theta, batch, fld = problem("H2O", 100, 6)
    model = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=True)
    params = model.unravel(dev(theta).requires_grad_(True))
    leaves = [params["params"]["cnf"]["net"]["VmapNN_0"][k][l] for k in ("last_layer", "layers_0", "layers_1") for l in ("bias", "kernel")]
    x, logp, score = neural_ode_score(params, dev(batch), model, 0.0, 1.0, 3, n_steps=4)
    assert x.shape == (200, 3) and logp.shape == (200, 1) and score.shape == (200, 3)
    wx, wl, ws = (np.random.default_rng(0).standard_normal(s) for s in ((200, 3), (200, 1), (200, 3)))
    loss = (x * dev(wx)).sum() + (logp * dev(wl)).sum() + (score * dev(ws)).sum()
    grads = torch.autograd.grad(loss, leaves)
    flat = torch.cat([g.reshape(-1) for g in grads]).cpu().numpy()
    _, stages = O.odeint_rk4(lambda y, t: fld.rhs(t, y, True), batch, 0.0, 1.0, 4, keep=True)
    _, tref = O.rk4_adjoint(fld, stages, 0.0, 1.0, np.concatenate([wx, wl, ws], 1), True)
    assert rel(flat, tref) < G_TOL
    z, lp = neural_ode(params, dev(batch[:, :4]), model, 0.0, 1.0, 3, n_steps=4)
    assert torch.equal(z, x.detach()) or np.abs(z.cpu().numpy() - x.detach().cpu().numpy()).max() < 1e-6
    # f.apply(params, t, states) == d/dt [x, logp]
    dy = model.apply(params, 0.25, dev(batch[:, :4])).cpu().numpy()
    assert np.abs(dy - fld.rhs(0.25, batch[:, :4], False)).max() < 2e-6 * max(1.0, np.abs(dy).max())
    # gradient through neural_ode (rows [x, logp]: the two-jet instantiation of the adjoint kernels, LiH.py / rho_rev paths)
    for path in (ops.PATH_DEFAULT, ops.PATH_FFMA):
        th = dev(theta).requires_grad_(True)
        y1, ck = ops.gnn_fwd(th.detach(), model.xyz_nuclei, model.z_one_hot, dev(batch[:, :4]), 200, 2, 2, 4, 0.0, 1.0, -1.0,
                             want_ckpt=True, want_act=True, path=path)
        ybar = np.concatenate([wx, wl], 1)
        g, yb0 = ops.gnn_bwd(th.detach(), model.xyz_nuclei, model.z_one_hot, ck, dev(ybar), 200, 2, 2, 4, 0.0, 1.0, -1.0,
                             want_ybar0=True, path=path)
        _, st2 = O.odeint_rk4(lambda y, t: fld.rhs(t, y, False), batch[:, :4], 0.0, 1.0, 4, keep=True)
        yb0_ref, t2 = O.rk4_adjoint(fld, st2, 0.0, 1.0, ybar, False)
        assert rel(g.cpu().numpy(), t2) < G_TOL, (path, rel(g.cpu().numpy(), t2))
        assert rel(yb0.cpu().numpy()[:, :4], yb0_ref[:, :4]) < G_TOL


def test_device_prior_batch_generator_and_rho_rev():
    """prior log-density/score and the batch generators on the device; rho_rev
    (H20_mol_ofdft_min_equiv_flows.py:132-137) = reverse no-score flow + prior."""
    Ne, atoms, z, coords = utils.coordinates("H2O")
    prior = utils.ProMolecularDensity(z, coords)
    rng = np.random.default_rng(13)
    x = (coords[rng.integers(0, 3, 500)] + rng.standard_normal((500, 3))).astype(np.float32)
    lp, sc = utils.promolecular_log_prob_score_device(prior, dev(x))
    lp_ref, sc_ref = O.promolecular_logp_score(x.astype(np.float64), coords, z)
    assert np.abs(lp.cpu().numpy() - lp_ref).max() < 2e-6 * np.abs(lp_ref).max()
    assert np.abs(sc.cpu().numpy() - sc_ref).max() < 2e-6 * max(1.0, np.abs(sc_ref).max())
    gen = utils.batch_generator_device(7, 20000, prior)
    b0 = next(gen); b1 = next(gen)
    assert b0.shape == (40000, 7) and not torch.equal(b0, b1)
    assert torch.equal(b0, next(utils.batch_generator_device(7, 20000, prior)))          # reproducible per (seed, call)
    bn = b0.cpu().numpy().astype(np.float64)
    lp_ref, sc_ref = O.promolecular_logp_score(bn[:, :3], coords, z)
    assert np.abs(bn[:, 3] - lp_ref[:, 0]).max() < 1e-5 and np.abs(bn[:, 4:] - sc_ref).max() < 1e-5
    w = z / z.sum()
    assert np.abs(bn[:, :3].mean(0) - (w[:, None] * coords).sum(0)).max() < 0.03          # E[x] = sum w_i R_i, sigma/sqrt(n) ~ 0.006
    second = (w[:, None] * (coords ** 2 + 1.0)).sum(0)
    assert np.abs((bn[:, :3] ** 2).mean(0) - second).max() < 0.06
    assert np.abs(np.corrcoef(bn[:20000, 0], bn[20000:, 0])[0, 1]) < 0.03                 # the two halves are independent draws
    b1d = next(utils.batche_generator_1D_device(3, 50000)).cpu().numpy()
    assert abs(b1d[:, 0].mean()) < 0.02 and abs(b1d[:, 0].std() - 1.0) < 0.02 and np.allclose(b1d[:, 2], -b1d[:, 0])
    assert np.allclose(b1d[:, 1], -0.5 * b1d[:, 0] ** 2 - 0.5 * np.log(2 * np.pi), atol=1e-6)
    # rho_rev: density of the flow evaluated at forward-flowed samples equals exp(logp) carried by the forward pass
    mol, nt, theta, batch, fld = problem("H2O", 100, 14)
    fwd = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=True)
    rev = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=False)
    params = fwd.unravel(dev(theta))
    x1, logp1, _ = neural_ode_score(params, dev(batch), fwd, 0.0, 1.0, 3, n_steps=32)
    rho = utils.rho_rev(params, x1, rev, prior, n_steps=32).cpu().numpy()
    assert np.abs(rho[:, 0] - np.exp(logp1.cpu().numpy()[:, 0])).max() < 2e-4 * np.exp(logp1.cpu().numpy()).max()
    # and against the float64 oracle with the same scheme
    frev = O.GNNField(fld.params, mol["coords"], mol["onehot"], bool_neg=False)
    x1n = x1.cpu().numpy().astype(np.float64)
    y0 = O.odeint_rk4(lambda yy, t: frev.rhs(t, yy, False), np.concatenate([x1n, np.zeros((200, 1))], 1), -1.0, 0.0, 32)
    rho_ref = np.exp(O.promolecular_logp_score(y0[:, :3], mol["coords"], mol["z"])[0][:, 0] - y0[:, 3])
    assert np.abs(rho[:, 0] - rho_ref).max() < 1e-5 * rho_ref.max()


def test_rmsprop_clip_step_and_short_training_run():
    """clip_by_global_norm(1.0) + rmsprop on the flat parameters, and a 20-step run of the whole
    loop (device batch generator -> fused step -> optimiser) that must lower the energy."""
    rng = np.random.default_rng(15)
    th = rng.standard_normal(4609); g = 3.0 * rng.standard_normal(4609); nu = np.abs(rng.standard_normal(4609)) * 0.1
    t_ref, nu_ref = O.rmsprop_clip_step(th, g, nu, 3e-4)
    t_d, g_d, nu_d = dev(th), dev(g), dev(nu)
    ops.rmsprop_clip_step(t_d, g_d, nu_d, 3e-4)
    assert np.abs(t_d.cpu().numpy() - t_ref).max() < 1e-6 and np.abs(nu_d.cpu().numpy() - nu_ref).max() < 1e-6
    Ne, atoms, z, coords = utils.coordinates("H2")
    prior = utils.ProMolecularDensity(z, coords)
    model = Gen_EqvFlow(3, (64, 64), coords, utils.jax_one_hot(z), bool_neg=True)
    theta = model.flat_theta(model.init(0)).contiguous()            # reference default init: identity flow
    nu_t = torch.zeros_like(theta)
    spec = ops.EnergySpec(Ne=Ne, d=3, kin="tf-w", hart="hartree", nuc="nuclei_potential", x="lda", coords=coords, z=z)
    est = EnergyEstimator(model, spec, n_steps=8)
    gen = utils.batch_generator_device(0, 4096, prior)
    energies = []
    for _ in range(20):
        sums, tbar, _ = est.step_flat(theta, next(gen))
        energies.append(sums[5].item())
        ops.rmsprop_clip_step(theta, tbar, nu_t, 3e-3)
    assert np.mean(energies[-5:]) < np.mean(energies[:5]) - 0.02, energies


def test_adam_clip_step_matches_optax_semantics():
    """OFDFT_NF.py:104-108: optax.chain(clip_by_global_norm(1.0), adam(lr)); three consecutive updates (bias correction)."""
    rng = np.random.default_rng(16)
    th = rng.standard_normal(8769); mu = np.zeros(8769); nu = np.zeros(8769)
    t_d, mu_d, nu_d = dev(th), dev(mu), dev(nu)
    th = t_d.cpu().numpy().astype(np.float64)
    for step in (1, 2, 3):
        g = (0.02 if step == 3 else 3.0) * rng.standard_normal(8769)          # clipped twice, unclipped once
        g = dev(g).cpu().numpy().astype(np.float64)
        th, mu, nu = O.adam_clip_step(th, g, mu, nu, step, 3e-4)
        ops.adam_clip_step(t_d, dev(g), mu_d, nu_d, step, 3e-4)
        assert np.abs(t_d.cpu().numpy() - th).max() < 2e-6 and np.abs(mu_d.cpu().numpy() - mu).max() < 1e-7
        assert np.abs(nu_d.cpu().numpy() - nu).max() < 1e-8


def test_device_batches_of_consecutive_calls_are_independent():
    """Philox streams of consecutive generator calls are disjoint: row-wise correlation between batch k and k+1 (and
    between the raw normal draws that make them) vanishes.  (Skip-ahead counts single outputs; a 3-D row consumes 5.)"""
    Ne, atoms, z, coords = utils.coordinates("H2")
    prior = utils.ProMolecularDensity(z, coords)
    gen = utils.batch_generator_device(11, 100000, prior)
    b = [next(gen).cpu().numpy().astype(np.float64) for _ in range(3)]
    for k in range(2):
        for c0 in range(3):
            for c1 in range(3):
                assert abs(np.corrcoef(b[k][:, c0], b[k + 1][:, c1])[0, 1]) < 0.02, (k, c0, c1)
    g1 = utils.batche_generator_1D_device(11, 100000)
    d = [next(g1).cpu().numpy().astype(np.float64)[:, 0] for _ in range(3)]
    for k in range(2):
        assert abs(np.corrcoef(d[k], d[k + 1])[0, 1]) < 0.02
        assert abs(np.corrcoef(d[k] ** 2, d[k + 1] ** 2)[0, 1]) < 0.02


# ---- 1-D tanh-MLP flow (LiH.py: CNF(1,(512,512,512)), tfw_1d + softc + attr) ------------------------------------

def _mlp_problem(feat, bs, seed):
    rng = np.random.default_rng(seed)
    p = O.init_params(rng, 2, feat, 1, last_scale=0.5)
    for w, b in p.hidden:
        b[:] = 0.1 * rng.standard_normal(b.shape)
    p.last[1][:] = 0.05
    theta = O.flatten_params(p).astype(np.float32).astype(np.float64)
    batch = O.batch_generator_1d(rng, bs).astype(np.float32).astype(np.float64)
    return theta, batch, O.MLPField(O.unflatten_params(theta, 2, feat), bool_neg=True)


def _mlp_slices(feat):
    H, L = feat[0], len(feat)
    sl = {"b_last": (0, 1), "w_last": (1, 1 + H), "b0": (1 + H, 1 + 2 * H), "w0": (1 + 2 * H, 1 + 4 * H)}
    o = 1 + 4 * H
    for l in range(1, L):
        sl[f"b{l}"] = (o, o + H); sl[f"w{l}"] = (o + H, o + H + H * H); o += H + H * H
    return sl


def test_energy_ema_matches_optax_semantics():
    rng = np.random.default_rng(3)
    ema = ops.EnergyEMA(n=6, decay=0.99, capacity=8)
    state = np.zeros(6)
    for t in range(1, 12):
        v = rng.standard_normal(6)
        ref, state = O.ema_update(v, state, t, 0.99)
        out = ema.update(torch.as_tensor(v, dtype=torch.float64, device="cuda")).cpu().numpy()
        assert np.allclose(out, ref, rtol=1e-10, atol=1e-14)   # fused multiply-add in the kernel vs two roundings in numpy
    assert np.allclose(ema.history[(11 - 1) % 8].cpu().numpy(), ref, rtol=1e-10, atol=1e-14)
    with pytest.raises(ValueError):
        ema.update(torch.zeros(5, dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("feat,bs", [((64, 64), 37), ((128, 128, 128), 20), ((512, 512, 512), 24), ((256, 256, 256, 256), 9)])
def test_mlp1d_forward_and_training_step_match_oracle(feat, bs):
    from ofdft_nflows_b200.cn_flows import Gen_CNFSimpleMLP, mlp_fwd
    theta, batch, fld = _mlp_problem(feat, bs, 21)
    n_steps = 4
    model = Gen_CNFSimpleMLP(1, feat, bool_neg=True)
    assert model.theta_size() == theta.size
    ref = O.odeint_rk4(lambda y, t: fld.rhs(t, y, True), batch, 0.0, 1.0, n_steps)
    for jets, w in ((3, 3), (2, 2), (1, 1)):
        y1 = mlp_fwd(dev(theta), dev(batch), model, 0.0, 1.0, jets, n_steps)[0].cpu().numpy()
        assert np.abs(y1[:, :w] - ref[:, :w]).max() < 1e-5 * max(1.0, np.abs(ref[:, :w]).max()), jets
    ospec = O.EnergySpec(Ne=2, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
    spec = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
    e_ref, terms_ref, g_ref, _ = O.loss_and_grad(fld, ospec, batch, n_steps=n_steps)
    # tcgen05 kernels (default), the fused FP32 integrator / adjoint sweep and the FP32 weight-gradient GEMM
    # ... each tensor-core variant with the forward's activation checkpoint (default) and with the recomputing adjoint
    for path, act in ((ops.PATH_DEFAULT, True), (ops.PATH_DEFAULT, False), (ops.PATH_FFMA, True),
                      (ops.PATH_FFMA | ops.PATH_FFMA_DW, True), (ops.PATH_FFMA_DW, True), (ops.PATH_FFMA_DW, False)):
        sums, tbar, _ = EnergyEstimator(model, spec, n_steps=n_steps, path=path, act_ckpt=act).step_flat(dev(theta), dev(batch))
        sums = sums.cpu().numpy(); g = tbar.cpu().numpy()
        for i in range(3):
            assert abs(sums[i] - terms_ref[i]) <= E_TOL * abs(terms_ref[i]), (path, act, i, sums[i], terms_ref[i])
        for name, (a, b) in _mlp_slices(feat).items():
            assert rel(g[a:b], g_ref[a:b]) <= G_TOL, (path, act, name, rel(g[a:b], g_ref[a:b]))
    sums, tbar, _ = EnergyEstimator(model, spec, n_steps=n_steps).step_flat(dev(theta), dev(batch))
    # f.apply(params, t, states) (cn_flows.py:173-182) and the score-augmented right-hand side
    params = model.unravel(dev(theta))
    for w in (2, 3):
        dy = model.apply(params, 0.3, dev(batch[:, :w])).cpu().numpy()
        ref_dy = fld.rhs(0.3, batch[:, :w] if w == 3 else batch[:, :2], w == 3)
        assert np.abs(dy - ref_dy[:, :w]).max() < 5e-6 * max(1.0, np.abs(ref_dy).max()), w
    # bitwise reproducible
    sums2, tbar2, _ = EnergyEstimator(model, spec, n_steps=n_steps).step_flat(dev(theta), dev(batch))
    assert torch.equal(tbar, tbar2)


def test_mlp1d_multi_tile_multi_chunk_batch():
    """B = 2*bs = 2300 rows: 18 row tiles (the last one ragged) and two 2048-row chunks of the adjoint (accumulate path),
    small widths to keep the oracle cheap; all-pairs soft-Coulomb mode of the estimator against the oracle as well."""
    from ofdft_nflows_b200.cn_flows import Gen_CNFSimpleMLP
    feat, bs, n_steps = (64, 64), 1150, 2
    theta, batch, fld = _mlp_problem(feat, bs, 23)
    model = Gen_CNFSimpleMLP(1, feat, bool_neg=True)
    for hart, mode in (("softc", "paired"), ("softc_allpairs", "allpairs")):
        ospec = O.EnergySpec(Ne=2, kin="tfw_1d", hart=hart, nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
        spec = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
        e_ref, terms_ref, g_ref, _ = O.loss_and_grad(fld, ospec, batch, n_steps=n_steps)
        est = EnergyEstimator(model, spec, n_steps=n_steps, hartree_mode=mode)
        sums, tbar, _ = est.step_flat(dev(theta), dev(batch))
        sums_n = sums.cpu().numpy(); g = tbar.cpu().numpy()
        for i in range(3):
            assert abs(sums_n[i] - terms_ref[i]) <= E_TOL * abs(terms_ref[i]), (mode, i, sums_n[i], terms_ref[i])
        assert abs(sums_n[5] - e_ref) <= E_TOL * abs(e_ref)
        for name, (a, b) in _mlp_slices(feat).items():
            assert rel(g[a:b], g_ref[a:b]) <= G_TOL, (mode, name, rel(g[a:b], g_ref[a:b]))
        _, tbar2, _ = est.step_flat(dev(theta), dev(batch))
        assert torch.equal(tbar, tbar2)


@pytest.mark.parametrize("bs", [500, 512, 700])
def test_mlp1d_activation_checkpoint_matches_recompute_at_bench_width(bs):
    """3 x 512 at batch sizes the oracle cannot afford: the adjoint that reads the forward's activation checkpoint (weight-gradient
    GEMM beside the sweep at B = 1000 / 1024 rows, after it at B = 1400; B = 1000 and 1400 are ragged in the 128-row tiles)
    against the recomputing adjoint, which the small-batch tests pin to the oracle."""
    from ofdft_nflows_b200.cn_flows import Gen_CNFSimpleMLP
    feat = (512, 512, 512)
    theta, batch, _ = _mlp_problem(feat, bs, 31)
    model = Gen_CNFSimpleMLP(1, feat, bool_neg=True)
    spec = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
    out = {}
    for act in (True, False):
        est = EnergyEstimator(model, spec, n_steps=2, act_ckpt=act)
        sums, tbar, _ = est.step_flat(dev(theta), dev(batch))
        _, tbar2, _ = est.step_flat(dev(theta), dev(batch))
        assert torch.equal(tbar, tbar2), act                  # bitwise reproducible, side stream or not
        out[act] = (sums.cpu().numpy(), tbar.cpu().numpy().astype(np.float64))
    assert np.array_equal(out[True][0], out[False][0])        # same forward
    for name, (a, b) in _mlp_slices(feat).items():
        assert rel(out[True][1][a:b], out[False][1][a:b]) <= 2e-5, (name, rel(out[True][1][a:b], out[False][1][a:b]))


def test_rho_rev_1d_and_compute_integral():
    """LiH.py:104-111: density of the 1-D flow on a grid (reverse no-score flow + N(0,1) prior) and its trapezoid integral;
    compute_integral (H20_mol_ofdft_min_equiv_flows.py:35-39) over a weighted grid."""
    from ofdft_nflows_b200.cn_flows import Gen_CNFSimpleMLP
    feat = (64, 64)
    theta, _, fld = _mlp_problem(feat, 8, 29)
    rev = Gen_CNFSimpleMLP(1, feat, bool_neg=False)
    frev = O.MLPField(fld.params, bool_neg=False)
    xg = np.linspace(-6.0, 6.0, 400)[:, None]
    params = rev.unravel(dev(theta))
    rho = utils.rho_rev_1d(params, dev(xg), rev, n_steps=16).cpu().numpy().astype(np.float64)
    y0 = O.odeint_rk4(lambda yy, t: frev.rhs(t, yy, False), np.concatenate([xg, np.zeros_like(xg)], 1), -1.0, 0.0, 16)
    rho_ref = np.exp(-0.5 * y0[:, 0] ** 2 - 0.5 * np.log(2 * np.pi) - y0[:, 1])
    assert np.abs(rho[:, 0] - rho_ref).max() < 1e-5 * rho_ref.max()
    w = np.full(400, xg[1, 0] - xg[0, 0]); w[[0, -1]] *= 0.5
    tot = utils.compute_integral(params, (dev(xg), w), lambda p, x: utils.rho_rev_1d(p, x, rev, n_steps=16), 2.0)
    assert abs(float(tot) - 2.0) < 2e-3            # the flow preserves the normalisation of the prior


def test_mlp1d_golden_reference_semantics_lih_config():
    """BASELINE.json configs[0]: LiH 1-D, 3x512 tanh MLP, TF+W kinetic, soft-Coulomb Hartree (adaptive Dopri5 fixture)."""
    import importlib.util
    from ofdft_nflows_b200.cn_flows import Gen_CNFSimpleMLP
    spec_ = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec_); spec_.loader.exec_module(mg)
    g = np.load(os.path.join(GOLD, "lih1d_512x3_bs48.npz"))
    feat = tuple(int(f) for f in g["features"])
    theta, _ = mg.theta_1d(int(g["seed"]), feat)
    model = Gen_CNFSimpleMLP(1, feat, bool_neg=True)
    spec = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
    sums, tbar, y1 = EnergyEstimator(model, spec, n_steps=16).step_flat(dev(theta), dev(g["batch"]))
    sums = sums.cpu().numpy(); grad = tbar.cpu().numpy().astype(np.float64)
    for i in range(3):
        assert abs(sums[i] - g["terms"][i]) <= E_TOL * abs(g["terms"][i]), (i, sums[i], g["terms"][i])
    assert abs(sums[5] - float(g["energy"])) <= E_TOL * abs(float(g["energy"]))
    assert np.allclose(mg.tensor_norms_1d(grad, feat), g["grad_norms"], rtol=G_TOL)
    assert rel(grad[::61], g["grad_sub"]) <= G_TOL
    # drop-in entry point with autograd
    params = model.unravel(dev(theta).requires_grad_(True))
    x, logp, score = neural_ode_score(params, dev(g["batch"]), model, 0.0, 1.0, 1, n_steps=16)
    assert np.abs(x.detach().cpu().numpy() - g["y1"][:, :1]).max() < 1e-4
    (x.sum() + logp.sum() + score.sum()).backward()


# ---- fixtures produced by the reference's own source (tests/golden/make_ref_golden.py, tests/_jaxstub) -----------------

def gnn_slices(n_in, feats):
    out, o = {}, 0
    out["b_last"] = (o, o + 1); o += 1
    out["w_last"] = (o, o + feats[-1]); o += feats[-1]
    prev = n_in
    for i, h in enumerate(feats):
        out[f"b{i}"] = (o, o + h); o += h
        out[f"W{i}"] = (o, o + prev * h); o += prev * h
        prev = h
    return out


REF_STEP_3D = [("ref_step_h2_bs40", 32), ("ref_step_lih_bs32", 32), ("ref_step_h2o_hgh_b88_vwn_bs48", 32),
               ("ref_step_c6h6_b88_pw92_bs24", 64), ("ref_step_h2o_zero_init_bs32", 16), ("ref_step_h2o_bench_bs512", 16),
               ("ref_step_h2o_L1_bs32", 32), ("ref_step_h2o_L3_bs32", 32)]


def _check_ref_step_3d(name, n_steps, path=ops.PATH_DEFAULT, act_ckpt=True):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    feats = tuple(int(v) for v in g["features"])
    kin, nuc, hart, x, c = (str(v) for v in g["names"])
    nt = g["onehot"].shape[1]
    model = Gen_EqvFlow(3, feats, g["coords"], g["onehot"], bool_neg=True)
    spec = ops.EnergySpec(Ne=float(g["Ne"]), d=3, kin=kin, hart=hart, nuc=nuc, x=x, c=c, coords=g["coords"], z=g["z"])
    est = EnergyEstimator(model, spec, n_steps=n_steps, path=path, act_ckpt=act_ckpt)
    sums, tbar, y1 = est.step_flat(dev(g["theta"]), dev(g["batch"]))
    sums = sums.cpu().numpy(); grad = tbar.cpu().numpy().astype(np.float64)
    ref_terms = g["terms"]
    scale = np.abs(ref_terms[:5]).max()
    for i in range(5):
        # a term that is a small difference of large per-sample values (e.g. the 1e-3 TF+W mean of benzene) is held to the
        # tolerance of the total it enters
        assert abs(sums[i] - ref_terms[i]) <= E_TOL * max(abs(ref_terms[i]), 1e-2 * scale), (i, sums[i], ref_terms[i])
    assert abs(sums[5] - float(g["energy"])) <= E_TOL * max(abs(float(g["energy"])), 1e-2 * scale)
    gref = g["grad"].astype(np.float64)
    for tname, (a, b) in gnn_slices(2 + nt, feats).items():
        if np.linalg.norm(gref[a:b]) == 0.0:
            assert np.linalg.norm(grad[a:b]) == 0.0, tname
        else:
            assert rel(grad[a:b], gref[a:b]) <= G_TOL, (tname, rel(grad[a:b], gref[a:b]))
    bs = g["batch"].shape[0] // 2
    y1 = y1.cpu().numpy()
    assert np.abs(y1[:bs] - g["y1"][:bs]).max() < 2e-5 * max(1.0, np.abs(g["y1"][:bs]).max())
    if "rev_x" in g.files:          # rho_rev through the package helper (H20_...py:132-137)
        rho = utils.rho_rev(model.unravel(dev(g["theta"])), dev(g["rev_x"]), Gen_EqvFlow(3, feats, g["coords"], g["onehot"], bool_neg=False),
                            utils.ProMolecularDensity(g["z"], g["coords"]), n_steps=n_steps)
        assert np.abs(rho.cpu().numpy().reshape(-1) / g["rev_rho"][:, 0] - 1.0).max() < 2e-5


@pytest.mark.parametrize("name,n_steps", REF_STEP_3D)
def test_reference_source_fixtures_3d(name, n_steps):
    """y(t1), every energy term and dE/dtheta as the REFERENCE'S OWN SOURCE computes them (unmodified functionals.py /
    jax_ode.py / equiv_flows.py run under the jax stand-in, adaptive Dopri5 1e-7, float64): the driver-default functional
    sets (dirac_b88_x_e + vwn_c_e / pw92_c_e), benzene, the --nn 1 / --nn 3 architectures and the benchmark's own
    parameters at RK4 x16.  North-star tolerances: energy terms 1e-5, gradient 1e-4 per tensor."""
    _check_ref_step_3d(name, n_steps)


@pytest.mark.parametrize("name,n_steps,path", [("ref_step_h2o_bench_bs512", 16, ops.PATH_DEEP), ("ref_step_c6h6_b88_pw92_bs24", 64, ops.PATH_DEEP),
                                               ("ref_step_h2_bs40", 32, ops.PATH_DEEP), ("ref_step_h2o_L3_bs32", 32, ops.PATH_FFMA),
                                               ("ref_step_h2o_L1_bs32", 32, ops.PATH_FFMA)])
def test_reference_source_fixtures_3d_other_kernels(name, n_steps, path):
    """The same reference-source fixtures through the kernels the default path does not take: the depth-generic tensor-core
    kernels (gnn_deep_tc.cu, the default of (64, 64, 64)) on the (64, 64) fixtures, and the plain FP32 general-depth kernels
    (gnn_general.cu) on the (64, 64, 64) / (64,) ones."""
    _check_ref_step_3d(name, n_steps, path=path, act_ckpt=False)


def test_three_layer_tensor_core_kernels_multi_tile_vs_fp32_kernels():
    """(64, 64, 64): ragged multi-tile batch (both row segments, x-only second half), tensor-core kernels against the
    independently written FP32 kernels; bitwise reproducible."""
    Ne, atoms, z, coords = utils.coordinates("H2O")
    oh = utils.jax_one_hot(z)
    model = Gen_EqvFlow(3, (64, 64, 64), coords, oh, bool_neg=True)
    g = torch.Generator().manual_seed(7)
    theta = model.flat_theta(model.init(3)).to("cuda:0")
    theta = theta + 0.05 * torch.randn(theta.numel(), generator=g).to("cuda:0")
    batch = next(utils.batch_generator(99, 300, utils.ProMolecularDensity(z, coords)))
    spec = ops.EnergySpec(Ne=Ne, d=3, kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=coords, z=z)
    out = []
    for path in (ops.PATH_DEFAULT, ops.PATH_DEFAULT, ops.PATH_FFMA):
        sums, tbar, y1 = EnergyEstimator(model, spec, n_steps=16, path=path).step_flat(theta, dev(batch))
        out.append((sums.cpu().numpy(), tbar.cpu().numpy().astype(np.float64), y1.cpu().numpy()))
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    assert np.abs(out[0][2] - out[2][2]).max() < 2e-5 * max(1.0, np.abs(out[2][2]).max())
    assert abs(out[0][0][5] - out[2][0][5]) <= E_TOL * abs(out[2][0][5])
    for tname, (a, b) in gnn_slices(2 + oh.shape[1], (64, 64, 64)).items():
        assert rel(out[0][1][a:b], out[2][1][a:b]) <= G_TOL, (tname, rel(out[0][1][a:b], out[2][1][a:b]))


@pytest.mark.parametrize("name,n_steps", [("ref_step_lih1d_512x3_bs24", 16), ("ref_step_mlp1d_64x2_bs40", 16)])
def test_reference_source_fixtures_1d(name, n_steps):
    """LiH.py defaults (tfw_1d / attr / softc / xc_1d, 3 x 512 MLP) from the reference's own cn_flows.py / functionals.py."""
    from ofdft_nflows_b200.cn_flows import Gen_CNFSimpleMLP
    g = np.load(os.path.join(GOLD, name + ".npz"))
    feats = tuple(int(v) for v in g["features"])
    kin, nuc, hart, c = (str(v) for v in g["names"])
    model = Gen_CNFSimpleMLP(1, feats, bool_neg=True)
    spec = ops.EnergySpec(Ne=float(g["Ne"]), d=1, kin=kin, hart=hart, nuc=nuc, x="", c=c, R=float(g["R"]), Z_alpha=float(g["Z_alpha"]),
                          Z_beta=float(g["Z_beta"]))
    sums, tbar, y1 = EnergyEstimator(model, spec, n_steps=n_steps).step_flat(dev(g["theta"]), dev(g["batch"]))
    sums = sums.cpu().numpy(); grad = tbar.cpu().numpy().astype(np.float64)
    for i in (0, 1, 2, 4):
        assert abs(sums[i] - g["terms"][i]) <= E_TOL * max(abs(g["terms"][i]), 1e-12), (i, sums[i], g["terms"][i])
    assert abs(sums[5] - float(g["energy"])) <= E_TOL * abs(float(g["energy"]))
    gref = g["grad"].astype(np.float64)
    for tname, (a, b) in _mlp_slices(feats).items():
        assert rel(grad[a:b], gref[a:b]) <= G_TOL, (tname, rel(grad[a:b], gref[a:b]))
    assert np.abs(y1.cpu().numpy() - g["y1"]).max() < 2e-5 * max(1.0, np.abs(g["y1"]).max())


# ---- full-size (BASELINE.json config) properties ---------------------------------------------------------------

def _full_problem(bs=65536):
    Ne, atoms, z, coords = utils.coordinates("H2O")
    oh = utils.jax_one_hot(z)
    rng = np.random.default_rng(4321)
    w1 = rng.standard_normal((5, 64)) / np.sqrt(5); w2 = rng.standard_normal((64, 64)) / 8; w3 = 0.5 * rng.standard_normal((64, 1)) / 8
    theta = np.concatenate([np.zeros(1), w3.ravel(), np.zeros(64), w1.ravel(), np.zeros(64), w2.ravel()]).astype(np.float32)
    batch = next(utils.batch_generator(1234, bs, utils.ProMolecularDensity(z, coords)))
    return Ne, z, coords, oh, theta, batch


def test_full_size_zero_init_is_identity_and_only_last_layer_has_gradient():
    Ne, z, coords, oh, theta, batch = _full_problem()
    theta0 = theta.copy(); theta0[:65] = 0.0                       # reference default init: last layer zero
    model = Gen_EqvFlow(3, (64, 64), coords, oh, bool_neg=True)
    spec = ops.EnergySpec(Ne=Ne, d=3, kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=coords, z=z)
    est = EnergyEstimator(model, spec, n_steps=16, skip_unused_jets=False)
    sums, tbar, y1 = est.step_flat(dev(theta0), dev(batch))
    assert torch.equal(y1, dev(batch))                              # g == 0  =>  x = z, logp/score untouched, bit-exact
    g = tbar.cpu().numpy()
    assert np.abs(g[65:]).max() == 0.0 and np.abs(g[:65]).max() > 0.0
    # energy of the un-flowed prior == plain functional means of the inputs
    ospec = O.EnergySpec(Ne=Ne, kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=coords, z=z)
    e_ref, terms_ref = O.energy_terms(ospec, batch.astype(np.float64), 3)
    assert abs(sums[5].item() - e_ref) <= 1e-6 * abs(e_ref)


def test_full_size_roundtrip_self_convergence_and_determinism():
    Ne, z, coords, oh, theta, batch = _full_problem()
    th, nuc, ohd, y = dev(theta), dev(coords), dev(oh), dev(batch)
    B = y.shape[0]
    # rev o fwd = id at full size (fwd: bool_neg=True on [0,1]; rev: bool_neg=False on [-1,0])
    y1 = ops.gnn_fwd(th, nuc, ohd, y[:, :4].contiguous(), B, 2, 2, 16, 0.0, 1.0, -1.0)[0]
    y0 = ops.gnn_fwd(th, nuc, ohd, y1, B, 2, 2, 16, -1.0, 0.0, 1.0)[0]
    # the reverse pass subtracts what the forward added to logp only if started from it: check x and the logp increment
    assert (y0[:, :3] - y[:, :3]).abs().max().item() < 2e-5
    assert ((y1[:, 3] - y[:, 3]) + (y0[:, 3] - y1[:, 3])).abs().max().item() < 2e-5
    # N vs 2N self-convergence of every energy term, and run-to-run bitwise determinism of energy and gradient
    model = Gen_EqvFlow(3, (64, 64), coords, oh, bool_neg=True)
    spec = ops.EnergySpec(Ne=Ne, d=3, kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=coords, z=z)
    s16, g16, _ = EnergyEstimator(model, spec, n_steps=16).step_flat(th, y)
    s32, g32, _ = EnergyEstimator(model, spec, n_steps=32).step_flat(th, y)
    s16b, g16b, _ = EnergyEstimator(model, spec, n_steps=16).step_flat(th, y)
    assert torch.equal(s16, s16b) and torch.equal(g16, g16b)
    a, b = s16.cpu().numpy(), s32.cpu().numpy()
    assert np.all(np.abs(a[:4] - b[:4]) <= E_TOL * np.abs(b[:4]))
    assert rel(g16.cpu().numpy(), g32.cpu().numpy()) <= G_TOL
    # skipping the unused second-half jets does not change the result; neither does recomputing the layer-2
    # activations in the backward pass instead of reading the HBM checkpoint
    s_full, g_full, _ = EnergyEstimator(model, spec, n_steps=16, skip_unused_jets=False).step_flat(th, y)
    assert np.allclose(s_full.cpu().numpy(), a, rtol=1e-12) and rel(g_full.cpu().numpy(), g16.cpu().numpy()) < 1e-6
    s_rc, g_rc, _ = EnergyEstimator(model, spec, n_steps=16, act_ckpt=False).step_flat(th, y)
    assert torch.equal(s_rc, s16) and rel(g_rc.cpu().numpy(), g16.cpu().numpy()) < 2e-5    # FFMA recompute vs tcgen05 path
    # sharding invariance: two half-batches (paired shards) average to the full-batch result (section 8e)
    from ofdft_nflows_b200.estimator import shard_batch
    est = EnergyEstimator(model, spec, n_steps=16)
    parts = [est.step_flat(th, shard_batch(y, r, 2).contiguous()) for r in range(2)]
    sm = (parts[0][0] + parts[1][0]) / 2
    gm = (parts[0][1] + parts[1][1]) / 2
    assert np.allclose(sm.cpu().numpy()[:6], a[:6], rtol=1e-9)
    assert rel(gm.cpu().numpy(), g16.cpu().numpy()) < 1e-5


def test_full_size_three_layer_network_properties():
    """BASELINE.json's headline batch (H2O, bs = 65 536) with the --nn 3 network on the depth-generic tensor-core kernels:
    zero-initialised last layer => identity flow and a gradient in the last layer only (bit-exact), rev o fwd = id, N vs 2N
    self-convergence of the energy terms and the gradient, run-to-run bitwise determinism, sharding invariance."""
    Ne, z, coords, oh, theta2, batch = _full_problem()
    rng = np.random.default_rng(99)
    theta = np.concatenate([theta2, np.zeros(64), (rng.standard_normal((64, 64)) / 8).ravel()]).astype(np.float32)
    feats = (64, 64, 64)
    model = Gen_EqvFlow(3, feats, coords, oh, bool_neg=True)
    spec = ops.EnergySpec(Ne=Ne, d=3, kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=coords, z=z)
    th, nuc, ohd, y = dev(theta), dev(coords), dev(oh), dev(batch)
    B = y.shape[0]
    theta0 = theta.copy(); theta0[:65] = 0.0
    sums0, tbar0, y10 = EnergyEstimator(model, spec, n_steps=16, skip_unused_jets=False).step_flat(dev(theta0), y)
    assert torch.equal(y10, y)
    g0 = tbar0.cpu().numpy()
    assert np.abs(g0[65:]).max() == 0.0 and np.abs(g0[:65]).max() > 0.0
    y1 = ops.gnn_fwd(th, nuc, ohd, y[:, :4].contiguous(), B, 2, 2, 16, 0.0, 1.0, -1.0)[0]
    y0 = ops.gnn_fwd(th, nuc, ohd, y1, B, 2, 2, 16, -1.0, 0.0, 1.0)[0]
    assert (y0[:, :3] - y[:, :3]).abs().max().item() < 2e-5
    assert ((y1[:, 3] - y[:, 3]) + (y0[:, 3] - y1[:, 3])).abs().max().item() < 2e-5
    s16, g16, _ = EnergyEstimator(model, spec, n_steps=16).step_flat(th, y)
    s32, g32, _ = EnergyEstimator(model, spec, n_steps=32).step_flat(th, y)
    s16b, g16b, _ = EnergyEstimator(model, spec, n_steps=16).step_flat(th, y)
    assert torch.equal(s16, s16b) and torch.equal(g16, g16b)
    a, b = s16.cpu().numpy(), s32.cpu().numpy()
    assert np.all(np.abs(a[:4] - b[:4]) <= E_TOL * np.abs(b[:4]))
    assert rel(g16.cpu().numpy(), g32.cpu().numpy()) <= G_TOL
    from ofdft_nflows_b200.estimator import shard_batch
    est = EnergyEstimator(model, spec, n_steps=16)
    parts = [est.step_flat(th, shard_batch(y, r, 2).contiguous()) for r in range(2)]
    assert np.allclose(((parts[0][0] + parts[1][0]) / 2).cpu().numpy()[:6], a[:6], rtol=1e-9)
    assert rel(((parts[0][1] + parts[1][1]) / 2).cpu().numpy(), g16.cpu().numpy()) < 1e-5


def test_tensor_core_diagnostics():
    """The two diagnostic entry points of the ABI (they ship in the library, so the suite launches them): a 3xTF32 GEMM through
    shared-memory and tensor-memory operands against float64, and the tcgen05.mma timing probe DESIGN.md quotes."""
    from ofdft_nflows_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    A = rng.standard_normal((128, 64)).astype(np.float32); W = rng.standard_normal((64, 64)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64)
    st = torch.cuda.current_stream().cuda_stream
    a_d, w_d = dev(A), dev(W)
    for passes, tol in ((3, 5e-6), (-1, 5e-6), (1, 2e-3)):      # SS 3xTF32, TS 3xTF32, plain TF32
        d = torch.full((128, 64), float("nan"), device="cuda")
        _lib.check(lib.ofdft_tc_gemm_test(st, a_d.data_ptr(), w_d.data_ptr(), d.data_ptr(), 128, passes), "tc_gemm_test")
        torch.cuda.synchronize()
        err = np.abs(d.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err <= tol, (passes, err)
    out = torch.zeros(1, dtype=torch.int64, device="cuda")
    clk = {}
    for ts in (0, 1):
        for N in (64, 128):
            _lib.check(lib.ofdft_tc_mma_rate(st, out.data_ptr(), N, ts, 1200, 1, 1), "tc_mma_rate")
            torch.cuda.synchronize()
            clk[(ts, N)] = int(out.item()) / 1200
    # more columns cost more, and an activation operand in tensor memory never costs more than one in shared memory
    assert 16 < clk[(1, 64)] < clk[(1, 128)] < 200 and clk[(1, 64)] <= clk[(0, 64)] and clk[(1, 128)] <= clk[(0, 128)], clk
    assert lib.ofdft_tc_mma_rate(st, out.data_ptr(), 100, 0, 10, 1, 1) == -1
