"""Hartree all-pairs throughput (BASELINE.json metric part 2: "Hartree pairs/s") -- run on the GPU box:
    python tests/bench_pairs.py [n] [out.json]
Cross-half all-pairs estimator: energy + forces on both point sets (two tile passes).  Reports unique pairs/s,
kernel pair-evaluations/s and the fraction of the measured SFU (rsqrt) peak, which bounds the forward pass."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofdft_nflows_b200 import ops  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    out = sys.argv[2] if len(sys.argv) > 2 else None
    rng = np.random.default_rng(0)
    x = torch.as_tensor(rng.standard_normal((n, 7)).astype(np.float32), device="cuda")
    xp = torch.as_tensor(rng.standard_normal((n, 7)).astype(np.float32), device="cuda")
    sfu = ops.microbench("sfu"); ffma = ops.microbench("ffma")
    res = {"n": n, "m": n, "sfu_peak_per_s": sfu, "ffma_peak_flops": ffma}
    for mode, grad in (("energy+forces(both sets)", True), ("energy only", False)):
        for _ in range(2):
            ops.hartree_allpairs(x, xp, 10.0, "hartree", want_grad=grad)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 3
        for _ in range(reps):
            es, _, _ = ops.hartree_allpairs(x, xp, 10.0, "hartree", want_grad=grad)
        b.record(); b.synchronize()
        t = a.elapsed_time(b) * 1e-3 / reps
        evals = (2 if grad else 1) * float(n) * n
        res[mode] = {"seconds": t, "unique_pairs_per_s": float(n) * n / t, "pair_evals_per_s": evals / t,
                     "frac_of_sfu_peak": evals / t / sfu, "energy": es.item()}
    print(json.dumps(res))
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
