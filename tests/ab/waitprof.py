"""Development tool: clocks thread 0 of every CTA of gnn_bwd_tc_kernel spends in each tensor-core wait (variant lib_waitprof.so)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
dev = torch.device("cuda", 0)
est, theta, batch, mol = bench._gnn_estimator("h2o", 65536, 0, dev, 16)
raw = ctypes.CDLL(os.environ["OFDFT_B200_LIB"])
buf = (ctypes.c_ulonglong * 8)()
for _ in range(2):
    est.step_flat(theta, batch)
raw.ofdft_debug_waitprof(buf, 1)
timers = []
est.step_flat(theta, batch, timers=timers)
torch.cuda.synchronize()
raw.ofdft_debug_waitprof(buf, 0)
ms = {n: a.elapsed_time(b) for n, a, b in timers}
print(ms)
names = ["barD before jet 0 (previous evaluation's dW2)", "barD before jet 1/2 (previous jet's dW2)", "barB before remainder of jet 1/2", "barB last (B1 results)", "barD at the end of a stage (dW2 fold)"]
tot_clk = ms["gnn_bwd"] * 1e-3 * 1.965e9 * 148
for i, n in enumerate(names):
    print(f"{n:55s} {buf[i] / 148 / 1.965e6:8.3f} ms per SM  ({100 * buf[i] / tot_clk:5.1f} % of the kernel)")
