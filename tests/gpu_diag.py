"""GPU diagnostics: run every kernel against the float64 oracle and dump the error table as JSON.
Usage (on the GPU box): python tests/gpu_diag.py [out.json]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ofdft_oracle as O  # noqa: E402
from ofdft_nflows_b200 import ops  # noqa: E402

dev = "cuda"
rep = {}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def maxabs(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max())


def t32(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)


def setup(molname, B, seed, last_scale=0.5):
    rng = np.random.default_rng(seed)
    mol = O.molecule(molname)
    nt = mol["onehot"].shape[1]
    p = O.init_params(rng, 2 + nt, (64, 64), 1, last_scale=last_scale)
    for w, b in p.hidden:
        b[:] = 0.1 * rng.standard_normal(b.shape)
    p.last[1][:] = 0.05
    theta = O.flatten_params(p).astype(np.float32).astype(np.float64)
    p = O.unflatten_params(theta, 2 + nt, (64, 64))
    batch = O.batch_generator_3d(rng, B // 2, mol["coords"], mol["z"]).astype(np.float32).astype(np.float64)
    return mol, nt, p, theta, batch


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/diag.json"
    print("device", torch.cuda.get_device_name(0))
    for molname, B in (("H2O", 300), ("H2", 256), ("C6H6", 130)):
        mol, nt, p, theta, batch = setup(molname, B, 7)
        fld = O.GNNField(p, mol["coords"], mol["onehot"], bool_neg=True)
        th, nuc, oh, y = t32(theta), t32(mol["coords"]), t32(mol["onehot"]), t32(batch)
        key = f"{molname}_B{B}"
        # ---- rhs ----
        for jets in (3, 2, 1):
            w = {3: 7, 2: 4, 1: 3}[jets]
            dy = ops.gnn_rhs(th, nuc, oh, y, jets, 0.3, -1.0).cpu().numpy()
            ref = fld.rhs(0.3, batch, jets == 3)
            rep[f"{key}_rhs_j{jets}_maxabs"] = maxabs(dy[:, :w], ref[:, :w])
        # ---- forward ----
        n_steps = 4
        ref1 = O.odeint_rk4(lambda yy, t: fld.rhs(t, yy, True), batch, 0.0, 1.0, n_steps)
        y1, ck = ops.gnn_fwd(th, nuc, oh, y, B, 3, 3, n_steps, 0.0, 1.0, -1.0, want_ckpt=True, want_act=True)
        y1n = y1.cpu().numpy()
        rep[f"{key}_fwd_j3_maxabs"] = maxabs(y1n, ref1)
        rep[f"{key}_fwd_j3_x"] = maxabs(y1n[:, :3], ref1[:, :3])
        rep[f"{key}_fwd_j3_logp"] = maxabs(y1n[:, 3], ref1[:, 3])
        rep[f"{key}_fwd_j3_s"] = maxabs(y1n[:, 4:], ref1[:, 4:])
        rep[f"{key}_fwd_j3_s_relmax"] = maxabs(y1n[:, 4:], ref1[:, 4:]) / float(np.abs(ref1[:, 4:]).max())
        na = B // 2
        y1m, ckm = ops.gnn_fwd(th, nuc, oh, y, na, 3, 1, n_steps, 0.0, 1.0, -1.0, want_ckpt=True, want_act=True)
        y1mn = y1m.cpu().numpy()
        rep[f"{key}_fwd_mixed_full"] = maxabs(y1mn[:na], ref1[:na])
        rep[f"{key}_fwd_mixed_xonly"] = maxabs(y1mn[na:, :3], ref1[na:, :3])
        y2 = ops.gnn_fwd(th, nuc, oh, y[:, :4].contiguous(), B, 2, 2, n_steps, 0.0, 1.0, -1.0)[0].cpu().numpy()
        rep[f"{key}_fwd_j2_maxabs"] = maxabs(y2, ref1[:, :4])
        # ---- functionals ----
        for nucname in ("nuclei_potential", "hgh"):
            ospec = O.EnergySpec(Ne=mol["Ne"], kin="tf-w", hart="hartree", nuc=nucname, x="lda",
                                 coords=mol["coords"], z=mol["z"])
            gspec = ops.EnergySpec(Ne=mol["Ne"], d=3, kin="tf-w", hart="hartree", nuc=nucname, x="lda",
                                   coords=mol["coords"], z=mol["z"])
            e_ref, terms_ref = O.energy_terms(ospec, ref1, 3)
            g_ref = O.energy_terms_vjp(ospec, ref1, 3)
            sums, yb = ops.functionals_fwd_bwd(gspec, t32(ref1))
            sums = sums.cpu().numpy()
            rep[f"{key}_func_{nucname}_terms_rel"] = [abs(sums[i] - terms_ref[i]) / max(abs(terms_ref[i]), 1e-30) for i in range(4)]
            rep[f"{key}_func_{nucname}_total_rel"] = abs(sums[5] - e_ref) / abs(e_ref)
            rep[f"{key}_func_{nucname}_ybar_rel"] = relerr(yb.cpu().numpy(), g_ref)
            integ = ops.functionals_integrands(gspec, t32(ref1)).cpu().numpy()
            rep[f"{key}_func_{nucname}_integrand_mean_rel"] = abs(integ[0].mean() - terms_ref[0]) / abs(terms_ref[0])
        # ---- backward (same cotangent on both sides) ----
        ospec = O.EnergySpec(Ne=mol["Ne"], kin="tf-w", hart="hartree", nuc="nuclei_potential", x="lda",
                             coords=mol["coords"], z=mol["z"])
        e_ref, terms_ref, tbar_ref, y1_ref = O.loss_and_grad(fld, ospec, batch, n_steps=n_steps)
        ybar1 = O.energy_terms_vjp(ospec, y1_ref, 3)
        _, stages = O.odeint_rk4(lambda yy, t: fld.rhs(t, yy, True), batch, 0.0, 1.0, n_steps, keep=True)
        ybar0_ref, _ = O.rk4_adjoint(fld, stages, 0.0, 1.0, ybar1, True)
        tb, yb0 = ops.gnn_bwd(th, nuc, oh, ck, t32(ybar1), B, 3, 3, n_steps, 0.0, 1.0, -1.0, want_ybar0=True)
        tbn = tb.cpu().numpy()
        rep[f"{key}_bwd_theta_rel"] = relerr(tbn, tbar_ref)
        L = {"b3": (0, 1), "w3": (1, 65), "b1": (65, 129), "W1": (129, 129 + 64 * (2 + nt)),
             "b2": (129 + 64 * (2 + nt), 193 + 64 * (2 + nt)), "W2": (193 + 64 * (2 + nt), tbn.size)}
        for nme, (a, b) in L.items():
            rep[f"{key}_bwd_{nme}_rel"] = relerr(tbn[a:b], tbar_ref[a:b])
        rep[f"{key}_bwd_ybar0_rel"] = relerr(yb0.cpu().numpy(), ybar0_ref)
        # mixed: second half x-only cotangent (exactly what the training step does)
        tbm, _ = ops.gnn_bwd(th, nuc, oh, ckm, t32(ybar1), na, 3, 1, n_steps, 0.0, 1.0, -1.0)
        rep[f"{key}_bwd_mixed_theta_rel"] = relerr(tbm.cpu().numpy(), tbar_ref)
        # determinism
        tb2, _ = ops.gnn_bwd(th, nuc, oh, ck, t32(ybar1), B, 3, 3, n_steps, 0.0, 1.0, -1.0)
        rep[f"{key}_bwd_bitwise_repeat"] = bool(torch.equal(tb, tb2))
    # ---- hartree all pairs ----
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1000, 7)).astype(np.float32); xp = rng.standard_normal((777, 7)).astype(np.float32)
    es, xb, xpb = ops.hartree_allpairs(t32(x), t32(xp), 10.0, "hartree")
    xd, xpd = x[:, :3].astype(np.float64), xp[:, :3].astype(np.float64)
    diff = xd[:, None, :] - xpd[None]
    zz = (diff ** 2).sum(-1) + 3e-5
    eref = 0.5 * 100.0 * (1 / np.sqrt(zz)).mean()
    gref = (-0.5 * 100.0 * diff / zz[..., None] ** 1.5) / (1000 * 777)
    rep["hartree_pairs_e_rel"] = abs(es.item() - eref) / eref
    rep["hartree_pairs_xbar_rel"] = relerr(xb.cpu().numpy(), gref.sum(1))
    rep["hartree_pairs_xpbar_rel"] = relerr(xpb.cpu().numpy(), -gref.sum(0))
    # ---- peaks ----
    rep["ffma_peak_tflops"] = ops.microbench("ffma") / 1e12
    rep["sfu_peak_tops"] = ops.microbench("sfu") / 1e12
    # ---- quick timing, H2O bs=65536 ----
    mol, nt, p, theta, batch = setup("H2O", 2 * 65536, 11)
    th, nuc, oh, y = t32(theta), t32(mol["coords"]), t32(mol["onehot"]), t32(batch)
    gspec = ops.EnergySpec(Ne=10, d=3, kin="tf-w", hart="hartree", nuc="hgh", x="lda", coords=mol["coords"], z=mol["z"])
    B = y.shape[0]
    for mode, (na, ja, jb) in {"ref_all_full": (B, 3, 3), "mixed": (B // 2, 3, 1)}.items():
        for it in range(2):
            torch.cuda.synchronize(); ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            y1, ck = ops.gnn_fwd(th, nuc, oh, y, na, ja, jb, 16, 0.0, 1.0, -1.0, want_ckpt=True)
            ev[1].record()
            sums, yb = ops.functionals_fwd_bwd(gspec, y1)
            ev[2].record()
            tb, _ = ops.gnn_bwd(th, nuc, oh, ck, yb, na, ja, jb, 16, 0.0, 1.0, -1.0)
            ev[3].record(); torch.cuda.synchronize()
        rep[f"time_H2O_bs65536_{mode}_ms"] = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        rep[f"energy_H2O_bs65536_{mode}"] = sums.cpu().numpy().tolist()
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        json.dump(rep, f, indent=1)
    for k, v in rep.items():
        print(k, v)


if __name__ == "__main__":
    main()
