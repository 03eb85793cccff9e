"""Development tool (GPU box): training-step time of the H2O workload by GNN depth (features (64,), (64,64), (64,64,64))."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from ofdft_nflows_b200 import ops
from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow
from ofdft_nflows_b200.estimator import EnergyEstimator
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mol, nuc, theta_np, batch_np = bench.make_problem("h2o", bs)
dev = torch.device("cuda:0")
batch = torch.as_tensor(batch_np, device=dev)
for feats in ((64,), (64, 64), (64, 64, 64)):
    model = Gen_EqvFlow(3, feats, mol["coords"], mol["onehot"], bool_neg=True, device=dev)
    spec = ops.EnergySpec(Ne=mol["Ne"], d=3, kin="tf-w", hart="hartree", nuc=nuc, x="lda", coords=mol["coords"], z=mol["z"])
    est = EnergyEstimator(model, spec, n_steps=16)
    g = torch.Generator().manual_seed(1)
    theta = model.flat_theta(model.init(1)).to(dev); theta = theta + 0.05 * torch.randn(theta.numel(), generator=g).to(dev)
    for _ in range(2):
        est.step_flat(theta, batch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        est.step_flat(theta, batch)
    torch.cuda.synchronize()
    print(feats, f"{(time.perf_counter() - t0) / 3 * 1e3:.2f} ms/step")
