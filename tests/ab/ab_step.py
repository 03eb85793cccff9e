"""A/B timing of kernel builds (development tool, run on the GPU box):
    python tests/ab/ab_step.py lib1.so lib2.so ...        # each in its own process via OFDFT_B200_LIB
Prints per library the forward / backward kernel times of the H2O bs=65536 step (CUDA events, 6 timed steps after 3
warm-ups, L2 flushed) and the relative deviation of energy / gradient from the first library."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(out):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import bench
    from ofdft_nflows_b200 import ops
    dev = torch.device("cuda", 0)
    wl = os.environ.get("AB_WORKLOAD", "h2o")
    bs = int(os.environ.get("AB_BS", "0")) or bench.WORKLOADS[wl][1]
    est, theta, batch, mol = bench._gnn_estimator(wl, bs, 0, dev, 16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        est.step_flat(theta, batch)
    torch.cuda.synchronize()
    acc = {}
    for _ in range(6):
        flush.fill_(1)
        timers = []
        sums, tbar, _ = est.step_flat(theta, batch, timers=timers)
        torch.cuda.synchronize()
        for name, a, b in timers:
            acc.setdefault(name, []).append(a.elapsed_time(b))
    res = {k: sum(v) / len(v) for k, v in acc.items()}
    res["min_bwd"] = min(acc["gnn_bwd"]); res["min_fwd"] = min(acc["gnn_fwd"])
    np.savez(out, sums=sums.cpu().numpy(), tbar=tbar.cpu().numpy())
    print("AB " + json.dumps(res), flush=True)


def main():
    import numpy as np
    libs = sys.argv[1:]
    ref = None
    for i, lib in enumerate(libs):
        env = dict(os.environ, OFDFT_B200_LIB=os.path.abspath(lib), AB_CHILD="1")
        out = f"/tmp/ab_{i}.npz"
        try:
            p = subprocess.run([sys.executable, __file__, out], env=env, capture_output=True, text=True, timeout=150)
        except subprocess.TimeoutExpired:
            print(lib, "TIMEOUT (hang?)", flush=True); continue
        line = [l for l in p.stdout.splitlines() if l.startswith("AB ")]
        if not line:
            print(lib, "FAILED", p.stdout[-500:], p.stderr[-1500:]); continue
        r = json.loads(line[0][3:])
        d = np.load(out)
        if ref is None:
            ref = d
        de = float(np.abs(d["sums"][:6] - ref["sums"][:6]).max() / np.abs(ref["sums"][:6]).max())
        dg = float(np.linalg.norm(d["tbar"] - ref["tbar"]) / np.linalg.norm(ref["tbar"]))
        print(f"{os.path.basename(lib):28s} fwd {r['gnn_fwd']:.3f} (min {r['min_fwd']:.3f})  bwd {r['gnn_bwd']:.3f} (min {r['min_bwd']:.3f})  "
              f"dE {de:.1e} dG {dg:.1e}", flush=True)


if __name__ == "__main__":
    if os.environ.get("AB_CHILD"):
        child(sys.argv[1])
    else:
        main()
