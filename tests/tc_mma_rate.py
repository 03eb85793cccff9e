"""Diagnostic (GPU box): clocks per tcgen05.mma kind::tf32 M128.N.K8 by N and A-operand source (python tests/tc_mma_rate.py)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofdft_nflows_b200 import _lib
lib = _lib.load()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
reps = 1200
for ts in (0, 1):
    for N in (32, 64, 128, 256):
        for nacc in (1, 2, 3, 4):
            if nacc * N > 448:
                continue
            _lib.check(lib.ofdft_tc_mma_rate(st, out.data_ptr(), N, ts, reps, 1, nacc), "rate")
            torch.cuda.synchronize()
            c = int(out.item())
            print(f"A from {'TMEM' if ts else 'smem'}  N={N:3d} accumulators={nacc}: {c / reps:7.1f} clocks per MMA "
                  f"({2 * 128 * N * 8 * reps / c:7.0f} flop/clk)")
