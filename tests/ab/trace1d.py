"""Development tool: timeline of the forward kernel chain of the 1-D flow from %globaltimer stamps (variant lib_trace.so)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from ofdft_nflows_b200 import _lib
from ofdft_nflows_b200.cn_flows import mlp_fwd, Gen_CNFSimpleMLP
from ofdft_nflows_b200 import ops
lib = _lib.load()
dev = torch.device("cuda", 0)
feat = (512, 512, 512)
model = Gen_CNFSimpleMLP(1, feat, bool_neg=True)
g = torch.Generator().manual_seed(0)
theta = (0.05 * torch.randn(model.theta_size(), generator=g)).to(dev)
batch = torch.randn(1024, 3, generator=g).to(dev)
for _ in range(3):
    mlp_fwd(theta, batch, model, 0.0, 1.0, ops.JETS_FULL, 16, want_ckpt=True)
torch.cuda.synchronize()
raw = ctypes.CDLL(os.environ["OFDFT_B200_LIB"])
n = 8192 * 8
buf = (ctypes.c_ulonglong * n)()
rc = raw.ofdft_debug_trace_dump(buf, n)
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 8).astype(np.int64)
a = a[a[:, 1] > 0]
a = a[np.argsort(a[:, 1])]
print("rc", rc, "records", len(a))
# a window in the middle of the last forward
w = a[-120:-60]
t0 = w[0, 1]
prev_end = None
for r in w:
    kind = "gemm" if r[0] == 0 else "glue"
    st = [(int(v) - t0) / 1000.0 if v else None for v in r[1:7]]
    if kind == "gemm":
        print(f"{kind} start {st[0]:8.2f} waitdone {st[1]:8.2f} first_full {st[2]:8.2f} loop_end {st[3]:8.2f} mma_done {st[4]:8.2f} exit {st[5]:8.2f}   | wait->full {st[2]-st[1]:.2f} loop {st[3]-st[2]:.2f} tail {st[5]-st[3]:.2f}")
    else:
        print(f"{kind} start {st[0]:8.2f} waitdone {st[1]:8.2f} mid {st[5]:8.2f}")
