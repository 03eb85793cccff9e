"""world_size-2 gloo test (CPU) of the multi-GPU host logic: identical sharding of both halves, one all-reduce of the
per-term sums and of the parameter gradient.  The CUDA operators are replaced by the float64 oracle inside the
spawned processes (tests may use oracle/); what is under test is EnergyEstimator's shard/reduce plumbing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ofdft_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _problem():
    rng = np.random.default_rng(5)
    mol = O.molecule("H2O")
    p = O.init_params(rng, 5, (64, 64), 1, last_scale=0.5)
    theta = O.flatten_params(p)
    batch = O.batch_generator_3d(rng, 8, mol["coords"], mol["z"])
    return mol, theta, batch


def _install_fake_ops(mol, n_steps):
    """oracle-backed stand-ins with the signatures of ofdft_nflows_b200.ops (CPU float64 tensors)."""
    from ofdft_nflows_b200 import ops
    state = {}

    def fld(theta):
        return O.GNNField(O.unflatten_params(theta.numpy().astype(np.float64), 5, (64, 64)), mol["coords"], mol["onehot"], True)

    def gnn_fwd(theta, nuclei, onehot, y_in, n_a, ja, jb, n_steps_, t0, t1, y0, want_ckpt=False, out=None, want_act=False, path=0):
        f = fld(theta)
        y1, stages = O.odeint_rk4(lambda y, t: f.rhs(t, y, True), y_in.numpy().astype(np.float64), t0, t1, n_steps_, keep=True)
        state["stages"] = stages
        return torch.as_tensor(y1), "ckpt"

    def functionals_fwd_bwd(spec, y1, want_bar=True):
        ospec = O.EnergySpec(Ne=spec.Ne, kin=spec.kin, hart=spec.hart, nuc=spec.nuc, x=spec.x, coords=mol["coords"], z=mol["z"])
        e, terms = O.energy_terms(ospec, y1.numpy(), 3)
        sums = torch.zeros(8, dtype=torch.float64)
        sums[:5] = torch.as_tensor(terms); sums[5] = e
        return sums, torch.as_tensor(O.energy_terms_vjp(ospec, y1.numpy(), 3))

    def gnn_bwd(theta, nuclei, onehot, ckpt, ybar1, n_a, ja, jb, n_steps_, t0, t1, y0, want_ybar0=False, path=0):
        _, tbar = O.rk4_adjoint(fld(theta), state["stages"], t0, t1, ybar1.numpy(), True)
        return torch.as_tensor(tbar), None

    def hartree_allpairs(x, xp, Ne, kind="hartree", want_grad=True, xbar_into=None, xpbar_into=None, norm_pairs=0.0, esum_into=None):
        e, gx, gxp = O.hartree_allpairs(x[:, :3].numpy().astype(np.float64), xp[:, :3].numpy().astype(np.float64), Ne)
        f = (x.shape[0] * xp.shape[0] / norm_pairs) if norm_pairs else 1.0         # block of a larger pair set
        if xbar_into is not None:
            xbar_into[:, :3] += f * torch.as_tensor(gx)
        if xpbar_into is not None:
            xpbar_into[:, :3] += f * torch.as_tensor(gxp)
        if esum_into is not None:
            esum_into += f * e
            return esum_into, None, None
        return torch.tensor([f * e], dtype=torch.float64), None, None

    ops.gnn_fwd, ops.functionals_fwd_bwd, ops.gnn_bwd = gnn_fwd, functionals_fwd_bwd, gnn_bwd
    ops.hartree_allpairs = hartree_allpairs
    ops.gnn_act_ckpt_bytes = lambda *a: 0


def _make_estimator(mol, n_steps, hartree_mode="paired"):
    from ofdft_nflows_b200 import ops
    from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow
    from ofdft_nflows_b200.estimator import EnergyEstimator
    model = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=True, device="cpu")
    spec = ops.EnergySpec(Ne=10, d=3, kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=mol["coords"], z=mol["z"])
    return EnergyEstimator(model, spec, n_steps=n_steps, hartree_mode=hartree_mode)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ofdft_nflows_b200.estimator import shard_batch
    mol, theta, batch = _problem()
    _install_fake_ops(mol, 4)
    local = torch.as_tensor(shard_batch(batch, rank, world))
    res = {}
    for mode in ("paired", "allpairs"):
        est = _make_estimator(mol, 4, mode)
        sums, tbar, _ = est.step_flat(torch.as_tensor(theta), local)
        res[f"sums_{mode}"] = sums.numpy(); res[f"tbar_{mode}"] = tbar.numpy()
    if rank == 0:
        np.savez(out, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process(tmp_path):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    mol, theta, batch = _problem()
    _install_fake_ops(mol, 4)
    for mode in ("paired", "allpairs"):          # allpairs: all-gather of coordinates + strip evaluation per rank
        est = _make_estimator(mol, 4, mode)
        sums, tbar, _ = est.step_flat(torch.as_tensor(theta), torch.as_tensor(batch))
        assert np.allclose(got[f"sums_{mode}"], sums.numpy(), rtol=1e-12, atol=1e-14), mode
        assert np.linalg.norm(got[f"tbar_{mode}"] - tbar.numpy()) < 1e-12 * np.linalg.norm(tbar.numpy()), mode
    # and the single-process all-pairs step equals the oracle's all-pairs loss
    fld = O.GNNField(O.unflatten_params(theta, 5, (64, 64)), mol["coords"], mol["onehot"], True)
    ospec = O.EnergySpec(Ne=10, kin="tf-w", hart="hartree_allpairs", nuc="hgh", x="lda", coords=mol["coords"], z=mol["z"])
    e_ref, _, g_ref, _ = O.loss_and_grad(fld, ospec, batch, n_steps=4)
    assert abs(sums.numpy()[5] - e_ref) < 1e-12 * abs(e_ref)
    assert np.linalg.norm(tbar.numpy() - g_ref) < 1e-11 * np.linalg.norm(g_ref)


def test_shard_batch_pairs_partners():
    from ofdft_nflows_b200.estimator import shard_batch
    b = np.arange(16 * 7, dtype=np.float32).reshape(16, 7)
    s1 = shard_batch(b, 1, 2)
    assert np.array_equal(s1[:4], b[4:8]) and np.array_equal(s1[4:], b[12:16])
    assert np.array_equal(np.concatenate([shard_batch(b, r, 4)[:2] for r in range(4)]), b[:8])


def test_chunked_backward_path_equals_single_launch_on_cpu_fakes():
    """Above the activation-checkpoint cap the estimator re-integrates and reverses row chunks (one per half here);
    with the oracle-backed stand-ins the summed gradient must equal the single-launch result to rounding."""
    from ofdft_nflows_b200 import ops
    mol, theta, batch = _problem()
    _install_fake_ops(mol, 4)
    est = _make_estimator(mol, 4)
    s0, g0, _ = est.step_flat(torch.as_tensor(theta), torch.as_tensor(batch))
    ops.gnn_act_ckpt_bytes = lambda B, n_a, ja, jb, Na, n: 1000 * B        # 16 rows x 1000 B > cap below
    est2 = _make_estimator(mol, 4)
    est2.act_ckpt_max_bytes = 4000
    s1, g1, _ = est2.step_flat(torch.as_tensor(theta), torch.as_tensor(batch))
    assert np.allclose(s0.numpy(), s1.numpy(), rtol=1e-13, atol=1e-15)       # chunk means re-weighted: rounding-level differences
    assert np.linalg.norm(g1.numpy() - g0.numpy()) < 1e-12 * np.linalg.norm(g0.numpy())
    # the all-pairs estimator couples every row with every partner: one launch without the activation checkpoint
    est3 = _make_estimator(mol, 4, "allpairs")
    s2, g2, _ = est3.step_flat(torch.as_tensor(theta), torch.as_tensor(batch))
    est4 = _make_estimator(mol, 4, "allpairs")
    est4.act_ckpt_max_bytes = 4000
    s3, g3, _ = est4.step_flat(torch.as_tensor(theta), torch.as_tensor(batch))
    assert np.allclose(s2.numpy(), s3.numpy(), rtol=1e-13) and np.linalg.norm(g3.numpy() - g2.numpy()) < 1e-12 * np.linalg.norm(g2.numpy())
