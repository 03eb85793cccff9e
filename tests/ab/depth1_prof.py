import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from ofdft_nflows_b200 import ops
from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow
from ofdft_nflows_b200.estimator import EnergyEstimator
L = int(sys.argv[1]); bs = int(sys.argv[2])
mol, nuc, theta_np, batch_np = bench.make_problem("h2o", bs)
dev = torch.device("cuda:0")
batch = torch.as_tensor(batch_np, device=dev)
model = Gen_EqvFlow(3, (64,) * L, mol["coords"], mol["onehot"], bool_neg=True, device=dev)
spec = ops.EnergySpec(Ne=mol["Ne"], d=3, kin="tf-w", hart="hartree", nuc=nuc, x="lda", coords=mol["coords"], z=mol["z"])
est = EnergyEstimator(model, spec, n_steps=16)
theta = model.flat_theta(model.init(1)).to(dev) + 0.01
for _ in range(2):
    est.step_flat(theta, batch)
torch.cuda.synchronize()
