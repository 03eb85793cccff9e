"""tcgen05 diagnostic (GPU box): python tests/tc_diag.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofdft_nflows_b200 import _lib  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(0)
A = rng.standard_normal((128, 64)).astype(np.float32)
W = (rng.standard_normal((64, 64)) / 8).astype(np.float32)
ref = A.astype(np.float64) @ W.astype(np.float64)
for M, passes in ((128, 1), (128, 3), (64, 1), (64, 3)):
    a = torch.as_tensor(A, device="cuda"); w = torch.as_tensor(W, device="cuda")
    d = torch.full((128, 64), float("nan"), device="cuda")
    rc = lib.ofdft_tc_gemm_test(torch.cuda.current_stream().cuda_stream, a.data_ptr(), w.data_ptr(), d.data_ptr(), M, passes)
    torch.cuda.synchronize()
    dn = d.cpu().numpy().astype(np.float64)
    if M == 128:
        err = np.abs(dn - ref).max() / np.abs(ref).max()
        print(f"M={M} passes={passes} rc={rc} max rel err {err:.3e}")
    else:
        # find where rows 0..63 of the product landed among the 128 dumped TMEM lanes
        r64 = ref[:64]
        lanes = []
        for i in range(64):
            dist = np.abs(dn - r64[i][None]).max(1)
            lanes.append(int(np.nanargmin(dist)))
        errs = [np.abs(dn[l] - r64[i]).max() for i, l in enumerate(lanes)]
        print(f"M={M} passes={passes} rc={rc} row->lane {lanes}\n   max err {max(errs):.3e}")

# MN-major stacked [hi;lo] mode: D = A[:64]^T-contraction over the first 64 rows ("trajectories")
a = torch.as_tensor(A, device="cuda"); w = torch.as_tensor(W, device="cuda")
d = torch.full((128, 64), float("nan"), device="cuda")
rc = lib.ofdft_tc_gemm_test(torch.cuda.current_stream().cuda_stream, a.data_ptr(), w.data_ptr(), d.data_ptr(), 64, 0)
torch.cuda.synchronize()
dn = d.cpu().numpy().astype(np.float64)
ref2 = A[:64].astype(np.float64).T @ W.astype(np.float64)       # [k (64)][j (64)] = sum_traj A[traj][k] W[traj][j]
got = dn[:64] + dn[64:]
print("MN-major stacked rc", rc, "max rel err", np.abs(got - ref2).max() / np.abs(ref2).max(), "hi-only err", np.abs(dn[:64] - ref2).max() / np.abs(ref2).max())
np.set_printoptions(precision=4, linewidth=200, suppress=True)
print("dn[:4,:6]\n", dn[:4, :6]); print("ref2[:4,:6]\n", ref2[:4, :6])
print("nan count", np.isnan(dn).sum(), "absmax", np.nanmax(np.abs(dn)))
# try to identify: correlate rows of dn[:64] with rows/cols of ref2
for name, cand in (("ref2", ref2), ("ref2.T", ref2.T), ("A[:64]@W (K-major style)", A[:64].astype(np.float64) @ W.astype(np.float64)), ("A[:64]@W.T", A[:64].astype(np.float64) @ W.astype(np.float64).T), ("A[:64].T@W.T", A[:64].astype(np.float64).T @ W.astype(np.float64).T)):
    print(name, "err", np.abs(got - cand).max() / np.abs(cand).max())

# TS mode (A operand in tensor memory) and the padded K=trajectory operands of the weight-gradient GEMM
for passes, name in ((-1, "TS-mode 3xTF32"), (-2, "K=traj stacked hi|lo")):
    a = torch.as_tensor(A, device="cuda"); w = torch.as_tensor(W, device="cuda")
    d = torch.full((128, 64), float("nan"), device="cuda")
    rc = lib.ofdft_tc_gemm_test(torch.cuda.current_stream().cuda_stream, a.data_ptr(), w.data_ptr(), d.data_ptr(), 128, passes)
    torch.cuda.synchronize()
    dn = d.cpu().numpy().astype(np.float64)
    if passes == -1:
        print(name, "rc", rc, "max rel err", np.abs(dn - ref).max() / np.abs(ref).max())
    else:
        P = 0.5 * A[:, (np.arange(64) + 17) % 64].astype(np.float64)
        ref3 = A.astype(np.float64).T @ P
        got = dn[:64] + dn[64:]
        print(name, "rc", rc, "max rel err", np.abs(got - ref3).max() / np.abs(ref3).max(), "hi rows only", np.abs(dn[:64] - ref3).max() / np.abs(ref3).max())
