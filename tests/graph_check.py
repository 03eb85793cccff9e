"""CUDA-graph capture check (GPU box): python tests/graph_check.py [workload]
Captures one fused training step (forward, functionals, adjoint) in a torch CUDA graph, replays it and compares with the eager
result; prints the per-step time of both."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ofdft_nflows_b200 import ops
from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow
from ofdft_nflows_b200.estimator import EnergyEstimator

wl = sys.argv[1] if len(sys.argv) > 1 else "lih1d"
molname, bs, nuc = bench.WORKLOADS[wl]
mol, nuc, theta_np, batch_np = bench.make_problem(wl, bs)
dev = torch.device("cuda:0")
if wl == "lih1d":
    from ofdft_nflows_b200.cn_flows import Gen_CNFSimpleMLP
    model = Gen_CNFSimpleMLP(1, bench.MLP_FEAT, bool_neg=True, device=dev)
    spec = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
else:
    model = Gen_EqvFlow(3, (64, 64), mol["coords"], mol["onehot"], bool_neg=True, device=dev)
    spec = ops.EnergySpec(Ne=mol["Ne"], d=3, kin="tf-w", hart="hartree", nuc=nuc, x="lda", coords=mol["coords"], z=mol["z"])
est = EnergyEstimator(model, spec, n_steps=16)
theta = torch.as_tensor(theta_np, device=dev); batch = torch.as_tensor(batch_np, device=dev)
for _ in range(3):
    s_e, g_e, _ = est.step_flat(theta, batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    s_e, g_e, _ = est.step_flat(theta, batch)
torch.cuda.synchronize()
t_eager = (time.perf_counter() - t0) / 10
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    est.step_flat(theta, batch)
torch.cuda.current_stream().wait_stream(side)
with torch.cuda.graph(g):
    s_g, g_g, _ = est.step_flat(theta, batch)
g.replay(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    g.replay()
torch.cuda.synchronize()
t_graph = (time.perf_counter() - t0) / 10
print(f"{wl}: eager {1e3*t_eager:.3f} ms/step, graph replay {1e3*t_graph:.3f} ms/step; "
      f"sums equal {torch.equal(s_e, s_g)}, grad equal {torch.equal(g_e, g_g)}")
