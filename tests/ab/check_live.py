import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from ofdft_nflows_b200 import ops
from ofdft_nflows_b200.cn_flows import Gen_CNFSimpleMLP, mlp_fwd
mol, nuc, theta_np, batch_np = bench.make_problem("lih1d", 512)
dev = torch.device("cuda:0")
model = Gen_CNFSimpleMLP(1, bench.MLP_FEAT, bool_neg=True, device=dev)
spec = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
theta = torch.as_tensor(theta_np, device=dev); batch = torch.as_tensor(batch_np, device=dev)
y1, ck = mlp_fwd(theta, batch, model, 0.0, 1.0, ops.JETS_FULL, 16, want_ckpt=True)
sums, ybar1 = ops.functionals_fwd_bwd(spec, y1)
print("ybar1 first half abs max per col", ybar1[:512].abs().amax(0).tolist())
print("ybar1 second half abs max per col", ybar1[512:].abs().amax(0).tolist())
