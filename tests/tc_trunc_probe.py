
import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
from ofdft_nflows_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(0)
A = rng.standard_normal((128, 64)).astype(np.float32); W = (rng.standard_normal((64, 64)) / 8).astype(np.float32)
def trunc(x): return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
def rna(x): return ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
a = torch.as_tensor(A, device="cuda"); w = torch.as_tensor(W, device="cuda"); d = torch.zeros((128, 64), device="cuda")
rc = lib.ofdft_tc_gemm_test(torch.cuda.current_stream().cuda_stream, a.data_ptr(), w.data_ptr(), d.data_ptr(), 128, -3)
torch.cuda.synchronize(); dn = d.cpu().numpy().astype(np.float64)
for name, f in (("trunc", trunc), ("rna", rna)):
    ref = f(A).astype(np.float64) @ f(W).astype(np.float64)
    print(name, rc, np.abs(dn - ref).max() / np.abs(ref).max())
