"""Development tool (GPU box): depth-generic tensor-core kernels (gnn_deep_tc.cu) against the FP32 kernels on the fixtures."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ofdft_nflows_b200 import ops
from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow
from ofdft_nflows_b200.estimator import EnergyEstimator
GOLD = os.path.join(ROOT, "tests", "golden")
dev = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float32), device="cuda:0")
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def slices(n_in, feats):
    out, o = {}, 0
    out["b_last"] = (o, o + 1); o += 1
    out["w_last"] = (o, o + feats[-1]); o += feats[-1]
    prev = n_in
    for i, h in enumerate(feats):
        out[f"b{i}"] = (o, o + h); o += h
        out[f"W{i}"] = (o, o + prev * h); o += prev * h
        prev = h
    return out


def run(name, n_steps, path_a, path_b, label):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    feats = tuple(int(v) for v in g["features"])
    kin, nuc, hart, x, c = (str(v) for v in g["names"])
    nt = g["onehot"].shape[1]
    model = Gen_EqvFlow(3, feats, g["coords"], g["onehot"], bool_neg=True)
    spec = ops.EnergySpec(Ne=float(g["Ne"]), d=3, kin=kin, hart=hart, nuc=nuc, x=x, c=c, coords=g["coords"], z=g["z"])
    res = []
    for path in (path_a, path_b):
        est = EnergyEstimator(model, spec, n_steps=n_steps, path=path, act_ckpt=False)
        sums, tbar, y1 = est.step_flat(dev(g["theta"]), dev(g["batch"]))
        torch.cuda.synchronize()
        res.append((sums.cpu().numpy(), tbar.cpu().numpy().astype(np.float64), y1.cpu().numpy()))
    (sa, ga, ya), (sb, gb, yb) = res
    gref = g["grad"].astype(np.float64)
    print(label, name, "y1 maxdiff", np.abs(ya - yb).max(), "energy", sa[5], sb[5], float(g["energy"]))
    for t, (a, b) in slices(2 + nt, feats).items():
        print(f"   {t:7s} a-vs-b {rel(ga[a:b], gb[a:b]):.2e}   a-vs-ref {rel(ga[a:b], gref[a:b]):.2e}   b-vs-ref {rel(gb[a:b], gref[a:b]):.2e}")


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "l2"):
    run("ref_step_h2o_bench_bs512", 16, ops.PATH_DEEP, ops.PATH_DEFAULT, "L2 deep vs tc")
    run("ref_step_c6h6_b88_pw92_bs24", 64, ops.PATH_DEEP, ops.PATH_DEFAULT, "L2 deep vs tc")
if which in ("all", "l3"):
    run("ref_step_h2o_L3_bs32", 32, ops.PATH_DEFAULT, ops.PATH_FFMA, "L3 deep vs general")
