#!/usr/bin/env python
"""bench.py -- training-step trajectories/s of the Monte-Carlo energy estimator (CNF fwd + energy + grad).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # restated reference (JAX unavailable) on the host cores

One "step" = one pass of the hot path over one synthetic batch: integrate 2*bs base samples through the CNF
(x, log rho, score for the first half; x for the second half), evaluate the OF-DFT functionals, back-propagate
to the flow parameters.  Unit: trajectories/s (one trajectory = one row of the ODE state).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = 64
WORKLOADS = {
    # name: (molecule, per-GPU bs, nuclear functional)       BASELINE.json configs[2] / [1] / [3] / [4]
    "h2o": ("H2O", 65536, "hgh"),
    "h2": ("H2", 4096, "nuclei_potential"),
    "lih": ("LiH", 131072, "nuclei_potential"),
    "c6h6": ("C6H6", 16384, "nuclei_potential"),
    "lih1d": ("LiH-1D", 512, "attraction"),                   # BASELINE.json configs[0] (LiH.py defaults)
}
MLP_FEAT = (512, 512, 512)


def algorithmic_flops_per_traj_eval(na: int, jets: int, nn: int = 2) -> float:
    """Mat-mul-only 2-jet count per (trajectory, RHS evaluation): Na * [(nn-1)*2*J*H^2 + 2*J*H + 2*H]
    (J = jets carried: 3 = x/logp/score, 1 = x only; nn hidden layers, 2 in every config of BASELINE.json)."""
    return na * ((nn - 1) * 2.0 * jets * H * H + 2.0 * jets * H + 2.0 * H)


def make_problem(workload: str, bs: int, seed_offset: int = 0, nn: int = 2):
    """Synthetic inputs of the named config: promolecular base samples + prior logp/score, LeCun-normal hidden
    layers (zero biases) and a non-zero last layer 0.5*N(0,1/64) (the reference's zero-init
    last layer would make the field trivial).  Uses only the package's own host-side input producers."""
    from ofdft_nflows_b200 import utils as U
    molname, _, nuc = WORKLOADS[workload]
    if workload == "lih1d":
        rng = np.random.default_rng(4321)
        Hm, L = MLP_FEAT[0], len(MLP_FEAT)
        parts = [np.zeros(1), 0.5 * rng.standard_normal(Hm) / math.sqrt(Hm), np.zeros(Hm), rng.standard_normal(2 * Hm) / math.sqrt(2)]
        for _ in range(L - 1):
            parts += [np.zeros(Hm), rng.standard_normal(Hm * Hm) / math.sqrt(Hm)]
        theta = np.concatenate(parts).astype(np.float32)
        batch = next(U.batche_generator_1D(1234 + seed_offset, bs))
        return {"Ne": 2, "coords": np.zeros((2, 3)), "z": np.array([3., 1.]), "onehot": None}, nuc, theta, batch
    Ne, atoms, z, coords = U.coordinates(molname)
    onehot = U.one_hot_encode(z) if molname == "C6H6" else U.jax_one_hot(z)     # OFDFT_NF.py:73 vs H20_...py:83-84
    mol = {"coords": coords, "z": z, "Ne": Ne, "onehot": onehot}
    nt = onehot.shape[1]
    rng = np.random.default_rng(4321)
    I = 2 + nt
    w1 = rng.standard_normal((I, H)) / math.sqrt(I)
    w2 = rng.standard_normal((H, H)) / math.sqrt(H)
    # keep the summed field benign for many-nucleus molecules (the displacement scales with Na)
    w3 = 0.5 * min(1.0, 3.0 / len(z)) * rng.standard_normal((H, 1)) / math.sqrt(H)
    # flat theta, ravel_pytree key order: last_layer.{bias,kernel}, layers_0.{bias,kernel}, layers_1.{bias,kernel}
    parts = [np.zeros(1), w3.ravel(), np.zeros(H), w1.ravel()]
    if nn >= 2:
        parts += [np.zeros(H), w2.ravel()]
    for _ in range(nn - 2):                                     # --nn 3 (OFDFT_NF.py:320-326): one more LeCun-normal hidden layer
        parts += [np.zeros(H), (rng.standard_normal((H, H)) / math.sqrt(H)).ravel()]
    theta = np.concatenate(parts).astype(np.float32)
    batch = next(U.batch_generator(1234 + seed_offset, bs, U.ProMolecularDensity(z, coords)))
    return mol, nuc, theta, batch


def measure_traffic(args, kernel_label: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel: this script re-runs the workload
    (two steps) in a child process under ncu and reads the second launch.  Returns None when ncu is not usable."""
    import csv, io, re, shutil, subprocess
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    # not when this process is itself being profiled (ncu exports its injection settings to the processes it launches)
    if ncu is None or any(k in os.environ for k in ("NV_NSIGHT_INJECTION_PORT_BASE", "NV_TPS_LAUNCH_TOKEN", "CUDA_INJECTION64_PATH")):
        return None
    name = re.split(r"[<\s+]", kernel_label.strip())[0]        # "gnn_bwd_tc_kernel<recompute>" -> gnn_bwd_tc_kernel
    child = [sys.executable, os.path.abspath(__file__), "--traffic-child", "--workload", args.workload, "--n-steps", str(args.n_steps),
             "--nn", str(args.nn), "--hartree", args.hartree]
    if args.bs:
        child += ["--bs", str(args.bs)]
    for flag, on in (("--recompute", args.recompute), ("--ffma", args.ffma), ("--full-jets", args.full_jets)):
        if on:
            child.append(flag)
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", f"regex:^{name}",
           "-s", "1", "-c", "1", "--csv"] + child
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=150, env=env).stdout
        rows = [r for r in csv.reader(io.StringIO(out[out.index('"ID"'):]))]
        hdr = rows[0]
        tot, kern = 0.0, None
        for r in rows[1:]:
            d = dict(zip(hdr, r))
            if d.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                unit = d["Metric Unit"].lower()
                scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[unit]
                tot += float(d["Metric Value"].replace(",", "")) * scale
                kern = d.get("Kernel Name")
        if kern is None:
            return None
        return {"bytes": tot, "source": f"measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of the second launch of {kern.split('(')[0]} "
                                        "in a child process of this command under ncu (two steps of the same workload; outside the timed region)"}
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                     "hw_power_brake_slowdown": 0x80}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(self.period)
        except Exception as e:   # never fail the benchmark because of the sampler
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_step(workload: str, bs: int, threads: int):
    """One fwd+grad step of the restated reference on the host: float64, generic nested autodiff RHS
    (vmap(jacrev) + value_and_grad + vjp), adaptive Dopri5 rtol=atol=1e-7, reverse-mode gradient.
    Returns (seconds, trajectories)."""
    import torch
    from oracle import ofdft_oracle as O, ref_torch as R
    torch.set_num_threads(threads)
    mol, nuc, theta, batch = make_problem(workload, bs)
    if workload == "lih1d":
        spec = O.EnergySpec(Ne=2, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
        model = R.GenCNFSimpleMLP(1, MLP_FEAT, bool_neg=True)
        t0 = time.perf_counter()
        R.value_and_grad_loss(model, theta.astype(np.float64), 2, MLP_FEAT, batch.astype(np.float64), spec, method="dopri5")
        return time.perf_counter() - t0, 2 * bs
    nt = mol["onehot"].shape[1]
    spec = O.EnergySpec(Ne=mol["Ne"], kin="tf-w", hart="hartree", nuc=nuc, x="lda", coords=mol["coords"], z=mol["z"])
    model = R.GenEqvFlow((H, H), mol["coords"], mol["onehot"], bool_neg=True)
    t0 = time.perf_counter()
    R.value_and_grad_loss(model, theta.astype(np.float64), 2 + nt, (H, H), batch.astype(np.float64), spec, method="dopri5")
    return time.perf_counter() - t0, 2 * bs


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  JAX/Flax are not installable in
    this image (no wheels, no network), so this times the reference-faithful restatement oracle/ref_torch.py
    on all host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = args.workload
    # size the sample so that (steps + warmup) steps take about two minutes
    cpu_reference_step(wl, 64, cores)                       # warm the thread pool / autograd caches
    t_probe, n_probe = cpu_reference_step(wl, 1024 if wl != "lih1d" else 128, cores)
    per_traj = t_probe / n_probe
    budget = 120.0 / max(1, args.steps + args.warmup)
    # capped at 2048: the reverse-mode graph of the adaptive solve grows with the batch (bs=8192 exceeds 62 GB of host RAM)
    bs = int(max(64 if wl == "lih1d" else 256, min(2048, budget / per_traj / 2)))
    if wl == "lih1d":
        bs = min(bs, 512)
    for _ in range(args.warmup):
        cpu_reference_step(wl, bs, cores)
    times = []
    for _ in range(args.steps):
        t, n = cpu_reference_step(wl, bs, cores)
        times.append(t)
    total = sum(times)
    value = n * args.steps / total
    molname, full_bs, nuc = WORKLOADS[wl]
    sample = (f"bs={bs} ({n} trajectories) of the {molname} bs={full_bs} workload per step; float64, nested-autodiff RHS, "
              f"adaptive Dopri5 rtol=atol=1e-7, reverse-mode gradient (oracle/ref_torch.py; JAX not installable here)")
    line = {
        "impl": "reference", "metric": "training-step trajectories/s (CNF fwd+grad+energy)", "value": value,
        "unit": "trajectories/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": ((f"LiH 1-D tanh-MLP flow {MLP_FEAT}, TF1D+W / attraction / soft-Coulomb, bs={full_bs} per GPU "
                                 f"({2 * full_bs} trajectories)") if wl == "lih1d" else
                                (f"{molname} 3-D equivariant flow (64,64), TF+W / {nuc} / paired Hartree / LDA, "
                                 f"bs={full_bs} per GPU ({2 * full_bs} trajectories)")) +
                               f"; reference arm: adaptive Dopri5 1e-7 in float64 on a bounded sample bs={bs} per step"},
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _launches_1d(n: int, L: int, ffma: bool, act: bool = False) -> int:
    """Kernel launches of one 1-D step: n = RHS evaluations.  Tensor-core path: prep (2), glue (n + 1) and (L - 1) layer
    GEMMs per evaluation forward; the adjoint adds head + tail kernels and the transposed GEMMs, plus either the recomputed
    forward GEMMs and two weight-gradient launches per layer at the end, or -- with the forward's activation checkpoint --
    (L - 1) weight-gradient launches per group of 4 evaluations on the second stream.  + 2 functionals, + reduces."""
    fwd = 2 + (n + 1) + n * (L - 1) if not ffma else 2
    if ffma:
        bwd = 2
    elif act:
        bwd = 2 + n + 1 + n * (L - 1)
    else:
        bwd = 2 + (n + 1) + n + 2 * n * (L - 1)
    # groups of 4 evaluations, the last 4 one by one; + one reduce per layer
    dw = ((n - 4 + 3) // 4 + min(n, 4) + 1) * (L - 1) if (act and not ffma) else 2 * (L - 1)
    return fwd + 2 + bwd + 1 + dw


def _timed(fn, reps: int, warm: int = 1) -> float:
    """Seconds per call of ``fn`` (device time, CUDA events on the current stream, after ``warm`` untimed calls)."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


def hartree_pairs_block(dev, n: int, ffma_peak: float, sfu_peak: float):
    """Second half of BASELINE.json's metric: Hartree pairs/s of the tiled all-pairs kernel (north-star item 2) on one GPU,
    n x n cross-set pairs resident in HBM: energy only, energy + forces on both point sets (two strip passes), and the 1-D
    soft-Coulomb variant.  Rooflines: one MUFU.RSQ per pair against the measured SFU rate (energy only: 7 FP32 + 1 MUFU
    per pair; 1-D: 3-6 FP32 + 1 MUFU) and 12 FP32 instructions per pair against the measured FP32 issue rate (3-D forces)."""
    import torch
    from ofdft_nflows_b200 import ops
    g = torch.Generator(device=dev); g.manual_seed(7)
    x = torch.randn((n, 7), device=dev, generator=g); xp = torch.randn((n, 7), device=dev, generator=g)
    x1 = torch.randn((n, 3), device=dev, generator=g); xp1 = torch.randn((n, 3), device=dev, generator=g)
    fp32_slots = ffma_peak / 2.0                      # FP32 instructions per second
    out = {"n": n, "m": n, "sfu_peak_per_s": sfu_peak, "fp32_instr_peak_per_s": fp32_slots,
           "peak_source": "ofdft_microbench_sfu / ofdft_microbench_ffma measured in this run"}
    t = _timed(lambda: ops.hartree_allpairs(x, xp, 10.0, "hartree", want_grad=False), 2)
    out["energy_only_3d"] = {"kernel": "hartree_pairs_kernel<3,false>", "seconds": t, "pairs_per_s": float(n) * n / t,
                             "roofline": {"bound": "sfu", "achieved": float(n) * n / t, "peak": sfu_peak, "unit": "rsqrt/s",
                                          "frac": float(n) * n / t / sfu_peak}}
    t = _timed(lambda: ops.hartree_allpairs(x, xp, 10.0, "hartree", want_grad=True), 2)
    ev = 2.0 * n * n
    bound = min(fp32_slots / 12.0, sfu_peak)
    out["energy_forces_3d"] = {"kernel": "hartree_pairs_kernel<3,true> x2 (forces on x, then on xp)", "seconds": t,
                               "pairs_per_s": float(n) * n / t, "pair_evals_per_s": ev / t,
                               "roofline": {"bound": "fp32", "achieved": ev / t, "peak": bound, "unit": "pair-evals/s",
                                            "frac": ev / t / bound, "fp32_instr_per_pair_eval": 12}}
    t = _timed(lambda: ops.hartree_allpairs(x1, xp1, 2.0, "softc", want_grad=True), 2)
    out["energy_forces_1d_softc"] = {"kernel": "hartree_pairs_kernel<1,true> x2", "seconds": t, "pairs_per_s": float(n) * n / t,
                                     "pair_evals_per_s": ev / t,
                                     "roofline": {"bound": "sfu", "achieved": ev / t, "peak": sfu_peak, "unit": "rsqrt/s",
                                                  "frac": ev / t / sfu_peak}}
    return out


def _gnn_estimator(workload: str, bs: int, rank: int, dev, n_steps: int, hartree_mode: str = "paired", **kw):
    import torch
    from ofdft_nflows_b200 import ops
    from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow
    from ofdft_nflows_b200.estimator import EnergyEstimator
    molname, _, nuc = WORKLOADS[workload]
    mol, nuc, theta_np, batch_np = make_problem(workload, bs, seed_offset=rank)
    model = Gen_EqvFlow(3, (H, H), mol["coords"], mol["onehot"], bool_neg=True, device=dev)
    spec = ops.EnergySpec(Ne=mol["Ne"], d=3, kin="tf-w", hart="hartree", nuc=nuc, x="lda", coords=mol["coords"], z=mol["z"])
    est = EnergyEstimator(model, spec, n_steps=n_steps, hartree_mode=hartree_mode, **kw)
    return est, torch.as_tensor(theta_np, device=dev), torch.as_tensor(batch_np, device=dev), mol


def _max_over_ranks(seconds: float, dev, world: int) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def strong_scaling_block(dev, rank: int, world: int, n_steps: int, global_bs: int, steps: int = 3):
    """BASELINE.json configs[3]: LiH 3-D equivariant flow, GLOBAL bs = 2^20 sharded over the N ranks (strong scaling: the total
    work is fixed), and the same global batch on one GPU alone (rank 0; the other ranks wait) for the speed-up."""
    import torch
    import torch.distributed as dist
    bs = global_bs // world
    est, theta, batch, _ = _gnn_estimator("lih", bs, rank, dev, n_steps)
    dist.barrier(); torch.cuda.synchronize()
    t_n = _max_over_ranks(_timed(lambda: est.step_flat(theta, batch), steps), dev, world)
    del batch
    torch.cuda.empty_cache()
    t_1 = 0.0
    if rank == 0:
        est1, theta1, batch1, _ = _gnn_estimator("lih", global_bs, 0, dev, n_steps)
        est1._dist = lambda: False                        # this rank alone: no collectives
        t_1 = _timed(lambda: est1.step_flat(theta1, batch1), 2)
        del batch1
        torch.cuda.empty_cache()
    dist.barrier()
    t_1 = _max_over_ranks(t_1, dev, world)
    return {"workload": f"LiH 3-D equivariant flow (64,64), global bs={global_bs} ({2 * global_bs} trajectories) sharded over {world} GPUs, "
                        f"RK4 x{n_steps}, paired Hartree", "scaling": "strong", "n_gpus": world, "per_gpu_bs": bs,
            "ms_per_step": 1e3 * t_n, "trajectories_per_s": 2.0 * global_bs / t_n,
            "one_gpu_same_global_batch": {"ms_per_step": 1e3 * t_1, "trajectories_per_s": 2.0 * global_bs / t_1,
                                          "note": "rank 0 alone; above the 64 GB activation-checkpoint cap the estimator runs "
                                                  "the step over row chunks (forward -> functionals -> adjoint per chunk)"},
            "speedup_vs_one_gpu": t_1 / t_n, "collectives_per_step": "2 all-reduces (8 doubles, n_theta floats); no data-path exchange"}


def allpairs_block_single(dev, n_steps: int, bs: int, steps: int = 2):
    """The all-pairs step on one GPU (no process group): same report as allpairs_block."""
    import torch
    est, theta, batch, _ = _gnn_estimator("c6h6", bs, 0, dev, n_steps, hartree_mode="allpairs")
    timers = []
    t = _timed(lambda: est.step_flat(theta, batch, timers=timers), steps)
    torch.cuda.synchronize()
    phase = {}
    for name, a, b in timers:
        phase.setdefault(name, []).append(a.elapsed_time(b))
    phase = {k: sum(v[1:]) / max(1, len(v) - 1) for k, v in phase.items()}
    pair_s = phase.get("hartree_pairs", 0.0) * 1e-3
    evals = 2.0 * bs * bs
    del batch
    torch.cuda.empty_cache()
    return {"workload": f"C6H6 3-D equivariant flow (64,64), all-pairs Hartree, bs={bs} on one GPU, RK4 x{n_steps}",
            "n_gpus": 1, "ms_per_step": 1e3 * t, "trajectories_per_s": 2.0 * bs / t, "pair_evals_per_step": evals,
            "pair_phase_ms": 1e3 * pair_s, "pair_evals_per_s": evals / pair_s if pair_s else None,
            "pair_evals_per_s_per_gpu": evals / pair_s if pair_s else None, "phase_ms_rank0": phase}


def allpairs_block(dev, rank: int, world: int, n_steps: int, bs: int, steps: int = 3):
    """BASELINE.json configs[4]: benzene-sized (12 nuclei) GNN field with the all-pairs Hartree estimator: coordinates
    all-gathered over NCCL (12 B per row; the local block is evaluated while the gather is in flight), every rank evaluates
    its (bs x N bs) strips with the tiled pair kernel (energy + forces on both halves), energy terms and theta-gradient
    all-reduced."""
    import torch
    import torch.distributed as dist
    est, theta, batch, _ = _gnn_estimator("c6h6", bs, rank, dev, n_steps, hartree_mode="allpairs")
    dist.barrier(); torch.cuda.synchronize()
    timers = []
    t = _max_over_ranks(_timed(lambda: est.step_flat(theta, batch, timers=timers), steps), dev, world)
    torch.cuda.synchronize()
    phase = {}
    for name, a, b in timers:
        phase.setdefault(name, []).append(a.elapsed_time(b))
    phase = {k: sum(v[1:]) / max(1, len(v) - 1) for k, v in phase.items()}          # skip the warm-up call
    pair_s = _max_over_ranks(phase.get("hartree_pairs", 0.0) * 1e-3, dev, world)
    # the two coordinate all-gathers on their own (12 B per row and half): what the local block has to cover
    xs = torch.empty((bs, 3), dtype=torch.float32, device=dev)
    x_all = torch.empty((world * bs, 3), dtype=torch.float32, device=dev)

    def gathers():
        dist.all_gather_into_tensor(x_all, xs)
        dist.all_gather_into_tensor(x_all, xs)
    gather_s = _max_over_ranks(_timed(gathers, 5, warm=2), dev, world)
    evals = 2.0 * bs * (bs * world) * world                 # both strips of every rank
    del batch
    torch.cuda.empty_cache()
    return {"workload": f"C6H6 3-D equivariant flow (64,64), all-pairs Hartree, bs={bs} per GPU ({bs * world} global), RK4 x{n_steps}",
            "n_gpus": world, "ms_per_step": 1e3 * t, "trajectories_per_s": 2.0 * bs * world / t,
            "pair_evals_per_step": evals, "pair_phase_ms": 1e3 * pair_s,
            "pair_evals_per_s": evals / pair_s if pair_s else None,
            "pair_evals_per_s_per_gpu": evals / pair_s / world if pair_s else None,
            "all_gather_ms_standalone": 1e3 * gather_s,
            "all_gather_note": "both coordinate all-gathers timed on their own; inside the step they are issued asynchronously and "
                               "the local block (phase local_block_and_gather_wait) is evaluated while they are in flight",
            "phase_ms_rank0": phase, "gathered_bytes_per_rank_per_step": 2 * 12 * bs * world}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="h2o", choices=sorted(WORKLOADS))
    ap.add_argument("--bs", type=int, default=0, help="per-GPU batch size (default: the workload's)")
    ap.add_argument("--n-steps", type=int, default=16, help="RK4 steps of the fixed-step integrator")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-jets", action="store_true", help="carry logp/score on the second half too (reference dataflow)")
    ap.add_argument("--hartree", default="paired", choices=["paired", "allpairs"],
                    help="allpairs: O(bs^2) cross-half Hartree estimator (coordinates all-gathered over NCCL, tiled pair kernel)")
    ap.add_argument("--recompute", action="store_true",
                    help="no HBM activation checkpoint: the adjoint recomputes the layer-2 jets (tensor cores; with --ffma: FP32 pipe)")
    ap.add_argument("--nn", type=int, default=2, choices=[1, 2, 3],
                    help="hidden layers of the GNN radial MLP (OFDFT_NF.py --nn: (64,) / (64,64) / (64,64,64)); BASELINE.json's configs use 2")
    ap.add_argument("--ffma", action="store_true", help="FP32-pipe kernels (ops.PATH_FFMA) instead of the tcgen05 kernels")
    ap.add_argument("--no-extra-blocks", action="store_true",
                    help="skip the hartree_pairs block (N=1) and the strong / allpairs blocks (N>1) of the default run")
    ap.add_argument("--pairs-n", type=int, default=1 << 20, help="points per set of the hartree_pairs block")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE.json configs[4]: all-pairs Hartree + functional sweep over GLOBAL bs = 16K .. 4M on the "
                         "benzene-sized GNN field at this GPU count (prints one JSON line with the list and exits)")
    ap.add_argument("--sweep-max", type=int, default=1 << 22, help="largest global batch of --sweep")
    ap.add_argument("--strong-global-bs", type=int, default=1 << 20, help="global batch of the strong-scaling block (N>1)")
    ap.add_argument("--allpairs-bs", type=int, default=16384,
                    help="per-GPU batch of the all-pairs block (N>1); the block also runs 8x this size (2^20 global on 8 GPUs)")
    ap.add_argument("--no-traffic-probe", action="store_true",
                    help="do not measure the dominant kernel's DRAM traffic (one extra launch of this workload in a child process under ncu)")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from ofdft_nflows_b200 import ops
    from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow
    from ofdft_nflows_b200.estimator import EnergyEstimator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL prints its version banner with a plain printf when the communicator is
        # created (NCCL_DEBUG=VERSION on the GPU boxes), so file descriptor 1 points at stderr while that happens
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    n_gpus = world

    if args.sweep:
        pts = []
        gbs = 1 << 14
        while gbs <= args.sweep_max:
            if world > 1:
                dist.barrier()
            blk = (allpairs_block(dev, rank, world, args.n_steps, gbs // world, steps=1 if gbs >= (1 << 20) else 2)
                   if world > 1 else allpairs_block_single(dev, args.n_steps, gbs, steps=1 if gbs >= (1 << 20) else 2))
            blk["global_bs"] = gbs
            pts.append(blk)
            gbs *= 4
        if rank == 0:
            print(json.dumps({"sweep": "C6H6 (12 nuclei) GNN field, all-pairs Hartree + functionals, global bs 16K-4M",
                              "n_gpus": world, "points": pts}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    molname, bs_default, nuc = WORKLOADS[args.workload]
    bs = args.bs or bs_default
    mol, nuc, theta_np, batch_np = make_problem(args.workload, bs, seed_offset=rank, nn=args.nn)
    B = 2 * bs
    na = mol["coords"].shape[0]
    is_1d = args.workload == "lih1d"
    if is_1d:
        from ofdft_nflows_b200.cn_flows import Gen_CNFSimpleMLP
        model = Gen_CNFSimpleMLP(1, MLP_FEAT, bool_neg=True, device=dev)
        spec = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1)
    else:
        model = Gen_EqvFlow(3, (H,) * args.nn, mol["coords"], mol["onehot"], bool_neg=True, device=dev)
        spec = ops.EnergySpec(Ne=mol["Ne"], d=3, kin="tf-w", hart="hartree", nuc=nuc, x="lda", coords=mol["coords"], z=mol["z"])
    path = (ops.PATH_FFMA | ops.PATH_FFMA_DW) if args.ffma else ops.PATH_DEFAULT
    est = EnergyEstimator(model, spec, n_steps=args.n_steps, skip_unused_jets=not args.full_jets, act_ckpt=not args.recompute,
                          hartree_mode=args.hartree, path=path)
    theta = torch.as_tensor(theta_np, device=dev)
    params = model.unravel(theta)
    batch = torch.as_tensor(batch_np, device=dev)
    if args.traffic_child:          # child of measure_traffic(): two steps of this workload for ncu to count DRAM bytes on, no output
        for _ in range(2):
            est.step_flat(theta, batch)
        torch.cuda.synchronize()
        return
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- roofline denominator: measured FP32 FFMA peak of this device (MEASURED_PEAKS.json has no FP32 figure) ----
    ffma_peak = ops.microbench("ffma")

    # ---- device-resident timed region ----
    for _ in range(args.warmup):
        est.step_flat(theta, batch)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ev, kern_ev = [], []
    for _ in range(args.steps):
        flush.fill_(1)                                           # flush L2 between timed iterations
        timers = []
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        sums, tbar, _ = est.step_flat(theta, batch, timers=timers)
        e.record()
        step_ev.append((s, e)); kern_ev.append(timers)
    barrier()
    clocks = sampler.stop()
    step_ms = [s.elapsed_time(e) for s, e in step_ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = total_ms.item() * 1e-3
    value = n_gpus * B * args.steps / total_s
    kern_ms = {}
    for timers in kern_ev:
        for name, s, e in timers:
            kern_ms.setdefault(name, []).append(s.elapsed_time(e))
    kern_avg = {k: sum(v) / len(v) for k, v in kern_ms.items()}

    # ---- end to end through the public API with HOST buffers ----
    host_batch = torch.as_tensor(batch_np).pin_memory()
    host_sums = torch.empty(8, dtype=torch.float64).pin_memory()
    host_grad = torch.empty(theta.numel(), dtype=torch.float32).pin_memory()

    def e2e_step():
        b = host_batch.to(dev, non_blocking=True)
        (energy, fv), grads = est.value_and_grad(params, b)
        host_sums.copy_(est.last_sums, non_blocking=True)
        host_grad.copy_(model.flat_theta(grads), non_blocking=True)

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    e2e_ev = []
    for _ in range(args.steps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); e2e_step(); e.record()
        e2e_ev.append((s, e))
    barrier()
    e2e_ms = torch.tensor([sum(s.elapsed_time(e) for s, e in e2e_ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * B * args.steps / (e2e_ms.item() * 1e-3)

    # ---- roofline of the dominant kernel (the discrete-adjoint backward) ----
    jets_b = 3 if args.full_jets else 1
    evals = 4 * args.n_steps
    if is_1d:
        Hm, L = MLP_FEAT[0], len(MLP_FEAT)
        fwd_flops = B * evals * ((L - 1) * 3 * 2.0 * Hm * Hm + 3 * 2.0 * Hm + 2 * 2.0 * Hm)    # 3.151 MFLOP
    else:
        fwd_flops = bs * evals * (algorithmic_flops_per_traj_eval(na, 3, args.nn) + algorithmic_flops_per_traj_eval(na, jets_b, args.nn))
    bwd_flops = 2.0 * fwd_flops
    kern_avg.setdefault("gnn_bwd", kern_avg.get("mlp_bwd", 0.0)); kern_avg.setdefault("gnn_fwd", kern_avg.get("mlp_fwd", 0.0))
    bwd_s = kern_avg["gnn_bwd"] * 1e-3
    fwd_s = kern_avg["gnn_fwd"] * 1e-3
    tensor_path = (not is_1d) and not args.ffma and args.nn != 1      # one hidden layer has no GEMM: plain FP32 kernels
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    mlp_tc = is_1d and not args.ffma
    if mlp_tc:
        # 1-D flow, tensor-core path: every hidden-layer GEMM launch executes (rows padded to 128) x H x H x 3 jets x
        # 3 passes; the adjoint adds the recompute and the transposed GEMMs plus the weight-gradient GEMMs.
        Hm, L = MLP_FEAT[0], len(MLP_FEAT)
        rows_pad = math.ceil(B / 128) * 128
        gemm = 2.0 * rows_pad * Hm * Hm * 3 * 3
        fwd_tensor_flops = evals * (L - 1) * gemm
        # weight gradients: 3 passes (hi.hi, lo.hi, hi.lo) over n_evals x B x 3 jets contraction rows (128 x 128 tiles)
        dw_flops = (L - 1) * 2.0 * Hm * Hm * (evals * math.ceil(B / 8) * 8 * 3) * 3
        act_1d = not args.recompute          # forward keeps the layer activations: the adjoint skips the recompute GEMMs
        bwd_tensor_flops = (1.0 if act_1d else 2.0) * evals * (L - 1) * gemm + dw_flops
        tf32_peak = 0.5 * peaks.get("bf16_tflops_sustained", 1400.0)
        peak3 = tf32_peak / 3.0
        roofline = {
            "bound": "tensor", "kernel": "mlp_layer_gemm_tc_kernel launches of the adjoint + mlp_dw_gemm_ts_kernel" +
                                         (" (on a second stream, beside the sweep)" if act_1d else ""),
            "achieved": bwd_flops / bwd_s / 1e12, "peak": peak3, "unit": "TFLOP/s", "achieved_is": "algorithmic fp32-equivalent flops / launch time",
            "frac": bwd_flops / bwd_s / 1e12 / peak3, "traffic": None,
            "peak_source": ("3xTF32 = 1/3 x 1/2 x bf16_tflops_sustained of MEASURED_PEAKS.json (kernel timed inside the step; of measured)"
                            if peaks else "1/3 x 1/2 x 1.4 PFLOP/s (B200_PROFILING.md fallback; of fallback)"),
            "frac_vs_burst_peak": bwd_flops / bwd_s / 1e12 / (0.5 * peaks.get("bf16_tflops", 1640.0) / 3.0),
            "algorithmic_flops_per_launch": bwd_flops,
            "executed_tensor_flops_per_launch": bwd_tensor_flops,
            "executed_tensor_frac_of_dense_tf32": bwd_tensor_flops / bwd_s / 1e12 / tf32_peak,
            "note": "3xTF32; one launch per hidden layer and RHS evaluation (a 128-row x 512-unit layer with three jets does "
                    "not fit on chip); the sweep keeps 64 of 148 SMs busy at bs=512, the weight-gradient GEMM the rest "
                    "-- see DESIGN.md section 3c",
            "activation_checkpoint": act_1d,
            "fp32_equivalent": {"algorithmic_flops_per_launch": bwd_flops, "achieved": bwd_flops / bwd_s / 1e12,
                                "ffma_peak": ffma_peak / 1e12, "frac_of_ffma_peak": bwd_flops / bwd_s / ffma_peak,
                                "peak_source": "ofdft_microbench_ffma measured in this run"},
            "kernel_ms": kern_avg,
            "fwd": {"kernel": "mlp_layer_gemm_tc_kernel + mlp_glue_tc_kernel", "achieved": fwd_flops / fwd_s / 1e12,
                    "frac": fwd_flops / fwd_s / 1e12 / peak3, "executed_tensor_frac_of_dense_tf32": fwd_tensor_flops / fwd_s / 1e12 / tf32_peak,
                    "fp32_equivalent_frac_of_ffma_peak": fwd_flops / fwd_s / ffma_peak},
            "step": {"algorithmic_flops": 3.0 * fwd_flops, "frac_of_ffma_peak": 3.0 * fwd_flops / (total_s / args.steps) / ffma_peak},
        }
    elif tensor_path:
        # tcgen05 kernels: the GEMM phases run as 3xTF32 on the tensor pipe.  Executed tensor work per launch =
        # MMAs issued x (2*128*64*8 flop); peak = dense TF32 = 1/2 of the MEASURED sustained bf16 figure (the kernel is
        # timed inside the step loop), else 1/2 of the profiling guide's 1.4 PFLOP/s fallback.
        mma = 2.0 * 128 * 64 * 8
        tiles_h, tiles_l = math.ceil(bs / 128), math.ceil(bs / 128)
        rc = 72 if args.recompute else 0          # tensor-core recompute of the layer-2 jets: one more 3xTF32 GEMM per jet
        per_h, per_l = (72 + 96 + rc), (24 + 32 + rc // 3) if jets_b == 1 else (72 + 96 + rc)
        fwd_h, fwd_l = 72, (24 if jets_b == 1 else 72)
        if args.nn == 3:
            # gnn_deep_tc.cu, per hidden GEMM layer and jet: 24 forward-recompute + 24 transposed MMAs and a 16-MMA N = 128
            # weight-gradient batch (2 units each); forward kernel: 24 per layer and jet
            per_h, per_l = 2 * 3 * (48 + 32), 2 * jets_b * (48 + 32)
            fwd_h, fwd_l = 2 * 72, 2 * 24 * jets_b
        bwd_tensor_flops = evals * na * mma * (tiles_h * per_h + tiles_l * per_l)
        fwd_tensor_flops = evals * na * mma * (tiles_h * fwd_h + tiles_l * fwd_l)
        tf32_peak = 0.5 * peaks.get("bf16_tflops_sustained", 1400.0)
        peak3 = tf32_peak / 3.0            # an fp32-accurate GEMM costs three TF32 passes
        roofline = {
            "bound": "tensor", "kernel": ("gnn_deep_bwd_tc_kernel<3>" if args.nn == 3 else
                                          "gnn_bwd_tc_kernel<recompute>" if args.recompute else "gnn_bwd_tc_kernel"),
            "achieved": bwd_flops / bwd_s / 1e12, "peak": peak3,
            "unit": "TFLOP/s", "achieved_is": "algorithmic fp32-equivalent flops / launch time", "frac": bwd_flops / bwd_s / 1e12 / peak3, "traffic": None,
            "peak_source": ("3xTF32 = 1/3 x 1/2 x bf16_tflops_sustained of MEASURED_PEAKS.json (kernel timed inside the step; of measured)"
                            if peaks else "1/3 x 1/2 x 1.4 PFLOP/s (B200_PROFILING.md fallback; of fallback)"),
            "frac_vs_burst_peak": bwd_flops / bwd_s / 1e12 / (0.5 * peaks.get("bf16_tflops", 1640.0) / 3.0),
            "algorithmic_flops_per_launch": bwd_flops,
            "executed_tensor_flops_per_launch": bwd_tensor_flops,
            "executed_tensor_frac_of_dense_tf32": bwd_tensor_flops / bwd_s / 1e12 / tf32_peak,
            "note": "3xTF32 (+1 stacked lo.lo pass in the weight-gradient GEMM): executed tensor flops are 3-4x the "
                    "algorithmic fp32 flops; the kernel is bounded by its FP32/SFU tanh-jet epilogues and operand staging, "
                    "see DESIGN.md section 3",
            "fp32_equivalent": {"algorithmic_flops_per_launch": bwd_flops, "achieved": bwd_flops / bwd_s / 1e12,
                                "ffma_peak": ffma_peak / 1e12, "frac_of_ffma_peak": bwd_flops / bwd_s / ffma_peak,
                                "peak_source": "ofdft_microbench_ffma measured in this run"},
            "kernel_ms": kern_avg,
            "backward_mode": ("hidden-layer jets recomputed on the tensor cores (no activation checkpoint)" if args.recompute or args.nn == 3
                              else "layer-2 activation checkpoint in HBM"),
            "fwd": {"kernel": "gnn_deep_fwd_tc_kernel<3>" if args.nn == 3 else "gnn_fwd_tc_kernel", "achieved": fwd_flops / fwd_s / 1e12, "frac": fwd_flops / fwd_s / 1e12 / peak3,
                    "executed_tensor_frac_of_dense_tf32": fwd_tensor_flops / fwd_s / 1e12 / tf32_peak,
                    "fp32_equivalent_frac_of_ffma_peak": fwd_flops / fwd_s / ffma_peak},
            "step": {"algorithmic_flops": 3.0 * fwd_flops, "frac_of_ffma_peak": 3.0 * fwd_flops / (total_s / args.steps) / ffma_peak},
        }
    else:
        achieved = bwd_flops / bwd_s / 1e12
        roofline = {
            "bound": "fp32", "kernel": ("mlp_bwd_sweep_kernel + mlp_dw_gemm_kernel" if is_1d else
                                        "gnn_bwd_kernel" if args.nn == 2 else f"gnn_general_bwd_kernel<{args.nn}>"),
            "achieved": achieved, "peak": ffma_peak / 1e12, "unit": "TFLOP/s",
            "frac": achieved / (ffma_peak / 1e12), "traffic": None,
            "peak_source": "ofdft_microbench_ffma measured in this run (FP32 FFMA pipe; MEASURED_PEAKS.json holds HBM/bf16 only)",
            "algorithmic_flops_per_launch": bwd_flops, "executed_over_algorithmic": 1.5 if args.recompute else 1.0,
            "backward_mode": "recompute" if args.recompute else "layer-2 activation checkpoint in HBM",
            "kernel_ms": kern_avg,
            "fwd": {"kernel": "mlp_fwd_kernel" if is_1d else ("gnn_fwd_kernel" if args.nn == 2 else f"gnn_general_fwd_kernel<{args.nn}>"),
                    "achieved": fwd_flops / fwd_s / 1e12,
                    "frac": fwd_flops / fwd_s / ffma_peak},
            "step": {"algorithmic_flops": 3.0 * fwd_flops, "frac": 3.0 * fwd_flops / (total_s / args.steps) / ffma_peak},
        }

    # DRAM traffic of the dominant kernel, measured live: the same workload runs two steps in a child process under ncu (two
    # counters, one pass, the second launch of the kernel); outside the timed region, never a timing.  Where ncu or the
    # counters are unavailable the figure falls back to the committed `ncu --set full` capture of the default workload.
    measured = None
    if rank == 0 and world == 1 and not args.no_traffic_probe and not is_1d:     # (1-D: the roofline is a chain of launches, not one kernel)
        measured = measure_traffic(args, roofline["kernel"])
    if measured is not None:
        roofline["traffic"] = measured["bytes"]
        roofline["traffic_source"] = measured["source"]
        roofline["algorithmic_bytes_without_activation_checkpoint"] = float(
            B * 28 * 2 + 4 * args.n_steps * 6 * 4 * B + (2 * math.ceil(bs / 128)) * theta.numel() * 4)
    try:
        if measured is None and args.workload == "h2o" and bs == 65536 and args.n_steps == 16 and not args.recompute and not args.full_jets and args.nn == 2:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_full_gnn.json" if tensor_path else "r01_ncu_full_gnn_ffma_kernels.json")))
            for k in prof:
                if ("gnn_bwd_tc_kernel" if tensor_path else "gnn_bwd_kernel") in k["kernel"]:
                    gb = lambda v: float(v.split()[0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[v.split()[1]]
                    roofline["traffic"] = gb(k["dram__bytes_read.sum"]) + gb(k["dram__bytes_write.sum"])
                    roofline["traffic_source"] = ("profiles/r02_ncu_full_gnn.json: dram__bytes_read + dram__bytes_write of one `ncu --set full` capture "
                                                  "of this command (bytes per launch; 12.9 GB of it is the layer-2 activation checkpoint); "
                                                  "algorithmic bytes without the checkpoint: state rows, stage checkpoint, tile partials")
                    roofline["algorithmic_bytes_without_activation_checkpoint"] = float(
                        B * 28 * 2 + 4 * args.n_steps * 6 * 4 * B + (2 * math.ceil(bs / 128)) * theta.numel() * 4)
    except Exception:
        pass

    # ---- the rest of BASELINE.json's metric in the same line: Hartree pairs/s (N = 1) and, where the path communicates,
    # the strong-scaling (configs[3]) and all-pairs (configs[4]) blocks (N > 1) -------------------------------------------
    extra = {}
    default_run = (args.workload == "h2o" and not args.bs and args.hartree == "paired" and not args.no_extra_blocks
                   and not args.recompute and not args.ffma and args.nn == 2)
    if default_run:
        del batch, flush
        torch.cuda.empty_cache()
        if world == 1:
            extra["hartree_pairs"] = hartree_pairs_block(dev, args.pairs_n, ffma_peak, ops.microbench("sfu"))
        else:
            extra["strong"] = strong_scaling_block(dev, rank, world, args.n_steps, args.strong_global_bs)
            extra["allpairs"] = [allpairs_block(dev, rank, world, args.n_steps, b) for b in (args.allpairs_bs, 8 * args.allpairs_bs)]

    line = None
    if rank == 0:
        line = {
            "metric": "training-step trajectories/s (CNF fwd+grad+energy)", "value": value, "unit": "trajectories/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (f"LiH 1-D tanh-MLP flow {MLP_FEAT}, TF1D+W / attraction / soft-Coulomb, bs={bs} per GPU "
                                    f"({B} trajectories), RK4 x{args.n_steps}") if is_1d else
                                   (f"{molname} 3-D equivariant flow ({','.join(['64'] * args.nn)}), TF+W / {nuc} / paired Hartree / LDA, "
                                    f"bs={bs} per GPU ({B} trajectories), RK4 x{args.n_steps}"),
                       "second_half_jets": "x only (denp/scorep are unused by the reference loss)" if not args.full_jets
                       else "full (reference dataflow)",
                       "l2": "flushed between timed iterations (256 MiB write)", "parallelism": f"batch-sharded x{n_gpus}"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": int(host_batch.numel() * 4),
                    "d2h_bytes_per_step": int(host_sums.numel() * 8 + host_grad.numel() * 4)},
            # 3-D: fwd, functionals x2, bwd, gradient reduce.  1-D (n = 4 n_steps evaluations, L-1 hidden GEMM layers):
            #   forward  tc: 2 weight preps + (n+1) glue + n (L-1) layer GEMMs          | ffma: prep + fused kernel
            #   adjoint  tc: 2 preps + (n+1) tail + n head + 2 n (L-1) layer GEMMs      | ffma: prep + fused sweep
            #   both: functionals x2, small-gradient reduce, (L-1) x (weight-gradient GEMM + reduce)
            # (--nn 3 and the tensor-core recompute mode add the weight-split kernel in front of the backward)
            "gpu_launches": ((5 + (1 if (tensor_path and (args.nn == 3 or args.recompute)) else 0)) if not is_1d
                             else _launches_1d(4 * args.n_steps, len(MLP_FEAT), args.ffma, not args.recompute)) * args.steps,
            "roofline": roofline,
            "energy_terms_last_step": [float(v) for v in sums.cpu().numpy()[:6]],
            "hartree": ({"mode": "paired"} if args.hartree == "paired" else
                        {"mode": "allpairs (energy + forces on both halves)",
                         "pair_evaluations_per_step": 2.0 * (bs * n_gpus) ** 2,
                         "pair_evaluations_per_s": 2.0 * (bs * n_gpus) ** 2 / (total_s / args.steps)}),
        }
        line.update(extra)
        if n_gpus == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cbs = 2048 if not is_1d else 256
            cpu_reference_step(args.workload, 64, cores)    # warm-up (thread pool, autograd caches)
            t, n = cpu_reference_step(args.workload, cbs, cores)
            line["cpu_baseline"] = {
                "value": n / t, "unit": "trajectories/s", "cores": cores, "kind": "port",
                "sample": f"bs={cbs} ({n} trajectories) of the same workload, one step; float64 nested-autodiff RHS + adaptive "
                          f"Dopri5 1e-7 + reverse-mode gradient (oracle/ref_torch.py, torch CPU; JAX not installable here)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
