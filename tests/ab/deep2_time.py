import os, sys, time
import torch
sys.path.insert(0, os.getcwd())
import bench
from ofdft_nflows_b200 import ops
from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow
from ofdft_nflows_b200.estimator import EnergyEstimator
mol, nuc, theta_np, batch_np = bench.make_problem("h2o", 65536)
dev = torch.device("cuda:0")
batch = torch.as_tensor(batch_np, device=dev); theta = torch.as_tensor(theta_np, device=dev)
model = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=True, device=dev)
spec = ops.EnergySpec(Ne=mol["Ne"], d=3, kin="tf-w", hart="hartree", nuc=nuc, x="lda", coords=mol["coords"], z=mol["z"])
for name, path, act in (("pipelined+ckpt", ops.PATH_DEFAULT, True), ("pipelined recompute", ops.PATH_DEFAULT, False), ("deep<2>", ops.PATH_DEEP, False)):
    est = EnergyEstimator(model, spec, n_steps=16, path=path, act_ckpt=act)
    for _ in range(2): est.step_flat(theta, batch)
    tm = []
    torch.cuda.synchronize()
    for _ in range(5):
        t = []; est.step_flat(theta, batch, timers=t); torch.cuda.synchronize(); tm.append({n: s.elapsed_time(e) for n, s, e in t})
    print(name, {k: round(sum(d[k] for d in tm) / len(tm), 3) for k in tm[0]})
