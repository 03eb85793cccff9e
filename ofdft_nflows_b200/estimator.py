"""Fused Monte-Carlo energy estimator: one training step's work (the path bench.py measures).

``EnergyEstimator.value_and_grad(params, batch)`` is the B200 counterpart of
``jax.value_and_grad(loss, has_aux=True)(params, batch)`` in the reference drivers
(H20_mol_ofdft_min_equiv_flows.py:150-178, LiH.py:118-146):

    forward integrator (rows [0,bs): x, logp, score + stage checkpoints; rows [bs,2bs): x only -- the reference
    never uses denp/scorep, H20_...py:154-156)  ->  fused functional reductions + cotangents  ->  discrete adjoint.

Three kernel launches + two tiny reductions on one stream; the result stays on the device until the caller reads it.
With ``torch.distributed`` initialised, every rank integrates its shard of the batch (both halves sharded
identically, so the paired Hartree estimator needs no coordinate exchange) and the per-term sums and the
parameter gradient are all-reduced once.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import torch

from . import ops
from .jax_ode import DEFAULT_N_STEPS


@dataclass
class F_values:
    """Per-term means (reference chex dataclass, H20_mol_ofdft_min_equiv_flows.py:43-49)."""
    energy: torch.Tensor
    kin: torch.Tensor
    vnuc: torch.Tensor
    hart: torch.Tensor
    xc: torch.Tensor


def shard_batch(batch, rank: int, world: int):
    """Rows of rank ``rank``: the same slice of BOTH halves, so that trajectory i and its Hartree partner i+bs land
    on the same GPU and the paired estimator needs no coordinate exchange.
    ``batch`` is (2*bs, w) with bs divisible by ``world``; returns (2*bs/world, w)."""
    bs = batch.shape[0] // 2
    if bs % world:
        raise ValueError(f"batch size {bs} is not divisible by the world size {world}")
    per = bs // world
    lo, hi = rank * per, (rank + 1) * per
    cat = torch.cat if isinstance(batch, torch.Tensor) else __import__("numpy").concatenate
    return cat([batch[lo:hi], batch[bs + lo:bs + hi]], 0)


class EnergyEstimator:
    def __init__(self, model: Any, spec: ops.EnergySpec, n_steps: int = DEFAULT_N_STEPS, t0: float = 0.0, t1: float = 1.0,
                 skip_unused_jets: bool = True, group: Optional[Any] = None, act_ckpt: bool = True,
                 act_ckpt_max_bytes: Optional[int] = None, hartree_mode: str = "paired", path: int = ops.PATH_DEFAULT):
        self.model, self.spec, self.n_steps, self.t0, self.t1 = model, spec, int(n_steps), float(t0), float(t1)
        # rows of the second half only supply xp: integrate them without the logp/score channels
        self.jets_b = ops.JETS_X if skip_unused_jets else ops.JETS_FULL
        self.group = group
        self.path = int(path)          # kernel-path selector (ops.PATH_*): tcgen05 kernels by default
        # "paired": the reference's row-paired O(bs) Hartree estimator (functionals.py:463-486);
        # "allpairs": the lower-variance O(bs^2) cross-half estimator of the same double integral (north star): the
        # second half's coordinates (and the first half's, for the partner forces) are all-gathered over NCCL and every
        # rank evaluates its bs/G x bs strip with the tiled pair kernel.
        if hartree_mode not in ("paired", "allpairs"):
            raise ValueError(hartree_mode)
        self.hartree_mode = hartree_mode
        if hartree_mode == "allpairs":
            import dataclasses
            self._hart_kind = spec.hart
            self.spec = dataclasses.replace(spec, hart="")        # local functionals without the paired Hartree term
        # keep the layer-2 activation jets of every RHS evaluation in HBM instead of recomputing them in the backward
        # pass (B200: 180 GB); falls back to recomputation above ``act_ckpt_max_bytes``
        # default cap: 64 GB, but never more than half of the device memory that is free at the first step
        self.act_ckpt = bool(act_ckpt)
        self.act_ckpt_max_bytes = None if act_ckpt_max_bytes is None else int(act_ckpt_max_bytes)
        self._is_gnn = hasattr(model, "xyz_nuclei")

    def _dist(self) -> bool:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def step_flat(self, theta: torch.Tensor, batch: torch.Tensor, timers: Optional[list] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(sums float64[8] = means [kin, vnuc, hart, x, c, total, 0, 0], theta_bar, y1) for the local shard,
        all-reduced (mean over ranks) when running distributed.  ``timers``: optional list that receives
        (kernel name, start event, end event) triples recorded on the current stream (bench.py roofline)."""
        m = self.model
        B = batch.shape[0]
        bs = B // 2

        def mark():
            if timers is None:
                return None
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            return ev

        if self._is_gnn:
            e0 = mark()
            Na = m.xyz_nuclei.shape[0]
            if self.act_ckpt_max_bytes is None:
                free = torch.cuda.mem_get_info(batch.device)[0] if batch.is_cuda else (64 << 30)
                self.act_ckpt_max_bytes = int(min(64 << 30, free // 2))
            act_bytes = ops.gnn_act_ckpt_bytes(B, bs, ops.JETS_FULL, self.jets_b, Na, self.n_steps)
            want_act = self.act_ckpt and act_bytes <= self.act_ckpt_max_bytes
            # activation checkpoint above the cap (many nuclei x large batch):
            #  * paired mode: every term is a mean over rows (i, i + bs), so row chunks go through forward -> functionals ->
            #    adjoint independently, each with a checkpoint that fits (same speed as below the cap);
            #  * all-pairs mode couples every row with every partner: one launch without the activation checkpoint, the
            #    adjoint recomputes the layer-2 jets on the tensor cores (~13 % slower than with the checkpoint)
            chunked = (self.act_ckpt and not want_act and B > 0 and len(getattr(m, "features", (64, 64))) == 2
                       and self.hartree_mode == "paired")
            if chunked:
                sums, tbar, y1 = self._gnn_step_row_chunks(theta, batch, bs, Na, timers)
                return self._reduce(sums, tbar, y1)
            y1, ckpt = ops.gnn_fwd(theta, m.xyz_nuclei, m.z_one_hot, batch, bs, ops.JETS_FULL, self.jets_b, self.n_steps,
                                   self.t0, self.t1, m.y0, want_ckpt=True, want_act=want_act, path=self.path)
            e1 = mark()
            sums, ybar1 = ops.functionals_fwd_bwd(self.spec, y1)
            if self.hartree_mode == "allpairs":
                self._allpairs_hartree(y1, ybar1, sums, timers)
            e2 = mark()
            tbar, _ = ops.gnn_bwd(theta, m.xyz_nuclei, m.z_one_hot, ckpt, ybar1, bs, ops.JETS_FULL, self.jets_b,
                                  self.n_steps, self.t0, self.t1, m.y0, path=self.path)
            e3 = mark()
            if timers is not None:
                timers += [("gnn_fwd", e0, e1), ("functionals", e1, e2), ("gnn_bwd", e2, e3)]
        else:
            from .cn_flows import mlp_fwd, mlp_bwd
            e0 = mark()
            y1, ckpt = mlp_fwd(theta, batch, m, self.t0, self.t1, ops.JETS_FULL, self.n_steps, want_ckpt=True, path=self.path,
                               act_ckpt=self.act_ckpt)
            e1 = mark()
            sums, ybar1 = ops.functionals_fwd_bwd(self.spec, y1)
            if self.hartree_mode == "allpairs":
                self._allpairs_hartree(y1, ybar1, sums, timers)      # soft-Coulomb pairs: coordinates are column 0
            e2 = mark()
            tbar, _ = mlp_bwd(theta, ckpt, ybar1, m, self.t0, self.t1, ops.JETS_FULL, self.n_steps, path=self.path)
            e3 = mark()
            if timers is not None:
                timers += [("mlp_fwd", e0, e1), ("functionals", e1, e2), ("mlp_bwd", e2, e3)]
        return self._reduce(sums, tbar, y1)

    def _reduce(self, sums, tbar, y1):
        if self._dist():
            import torch.distributed as dist
            ws = dist.get_world_size(self.group)
            # one latency-bound collective each: 8 doubles and n_theta floats (mean of equally sized shards)
            dist.all_reduce(sums, group=self.group)
            dist.all_reduce(tbar, group=self.group)
            sums /= ws
            tbar /= ws
        return sums, tbar, y1

    def _gnn_step_row_chunks(self, theta, batch, bs: int, Na: int, timers: Optional[list] = None):
        """The whole step over row chunks [lo, hi) of BOTH halves (a trajectory and its Hartree partner stay together): the
        per-term means and the parameter gradient are sums of the chunks' contributions weighted by their share of the rows."""
        m = self.model
        per_pair = max(1, ops.gnn_act_ckpt_bytes(256, 128, ops.JETS_FULL, self.jets_b, Na, self.n_steps) // 128)
        rows = max(128, (self.act_ckpt_max_bytes // per_pair) // 128 * 128)
        sums = torch.zeros(8, dtype=torch.float64, device=batch.device)
        tbar = torch.zeros_like(theta)
        y1 = torch.empty_like(batch)
        for lo in range(0, bs, rows):
            hi = min(bs, lo + rows)
            n = hi - lo
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timers is not None else None
            if ev: ev[0].record()
            sub = torch.cat([batch[lo:hi], batch[bs + lo:bs + hi]], 0)
            y1c, ck = ops.gnn_fwd(theta, m.xyz_nuclei, m.z_one_hot, sub, n, ops.JETS_FULL, self.jets_b, self.n_steps,
                                  self.t0, self.t1, m.y0, want_ckpt=True, want_act=True, path=self.path)
            if ev: ev[1].record()
            sc, ybar = ops.functionals_fwd_bwd(self.spec, y1c)
            w = n / bs
            sums += sc * w
            ybar *= w
            if ev: ev[2].record()
            g, _ = ops.gnn_bwd(theta, m.xyz_nuclei, m.z_one_hot, ck, ybar, n, ops.JETS_FULL, self.jets_b, self.n_steps,
                               self.t0, self.t1, m.y0, path=self.path)
            tbar += g
            if ev:
                ev[3].record()
                timers += [("gnn_fwd", ev[0], ev[1]), ("functionals", ev[1], ev[2]), ("gnn_bwd", ev[2], ev[3])]
            y1[lo:hi] = y1c[:n]
            y1[bs + lo:bs + hi] = y1c[n:]
            del ck, y1c, ybar, sub
        return sums, tbar, y1

    def _allpairs_hartree(self, y1: torch.Tensor, ybar1: torch.Tensor, sums: torch.Tensor, timers: Optional[list] = None) -> None:
        """Adds the all-pairs Hartree energy of this rank's strips to ``sums`` and its forces to ``ybar1`` (in place).

        Distributed: only the coordinates travel -- (bs, d) float32 per half, 12 B per row in 3-D -- in two asynchronous
        NCCL all-gathers; the local block (this rank's x against its own xp: energy and forces on both halves) is evaluated
        while they are in flight, then the strips against the remote rows.  Every pair of the global cross set is visited by
        exactly one rank for the energy and for each of its two forces."""
        bs = y1.shape[0] // 2
        d = 3 if ops.HART[self._hart_kind] == 1 else 1
        Ne = self.spec.Ne
        x_loc, xp_loc = y1[:bs], y1[bs:]                       # contiguous row blocks; coordinates are the first d columns
        esum = torch.zeros(1, dtype=torch.float64, device=y1.device)

        def mark():
            if timers is None:
                return None
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            return ev

        e0 = mark()
        if not self._dist():
            ops.hartree_allpairs(x_loc, xp_loc, Ne, self._hart_kind, xbar_into=ybar1[:bs], xpbar_into=ybar1[bs:], esum_into=esum)
        else:
            import torch.distributed as dist
            ws, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            xs = x_loc[:, :d].contiguous()
            xps = xp_loc[:, :d].contiguous()
            x_all = torch.empty((ws * bs, d), dtype=y1.dtype, device=y1.device)
            xp_all = torch.empty_like(x_all)
            w1 = dist.all_gather_into_tensor(xp_all, xps, group=self.group, async_op=True)
            w2 = dist.all_gather_into_tensor(x_all, xs, group=self.group, async_op=True)
            norm = float(bs) * float(bs) * ws                  # pairs of this rank's strip
            ops.hartree_allpairs(x_loc, xp_loc, Ne, self._hart_kind, xbar_into=ybar1[:bs], xpbar_into=ybar1[bs:], esum_into=esum,
                                 norm_pairs=norm)
            w1.wait(); w2.wait()
            if timers is not None:
                timers.append(("local_block_and_gather_wait", e0, mark()))
            for lo, hi in ((0, rank), (rank + 1, ws)):
                if hi > lo:
                    # (local x) x (remote xp): energy + forces on the local x rows; (local xp) x (remote x): forces on xp
                    ops.hartree_allpairs(x_loc, xp_all[lo * bs:hi * bs], Ne, self._hart_kind, xbar_into=ybar1[:bs], esum_into=esum,
                                         norm_pairs=norm)
                    scratch_e = torch.zeros(1, dtype=torch.float64, device=y1.device)
                    ops.hartree_allpairs(xp_loc, x_all[lo * bs:hi * bs], Ne, self._hart_kind, xbar_into=ybar1[bs:],
                                         esum_into=scratch_e, norm_pairs=norm)
        if timers is not None:
            timers.append(("hartree_pairs", e0, mark()))
        sums[2:3] += esum
        sums[5:6] += esum

    def value_and_grad(self, params: Dict[str, Any], batch: torch.Tensor):
        """-> ((energy, F_values), grads) with ``grads`` shaped like ``params``."""
        theta = self.model.flat_theta(params).detach()
        sums, tbar, _ = self.step_flat(theta, batch)
        self.last_sums = sums
        fv = F_values(energy=sums[5], kin=sums[0], vnuc=sums[1], hart=sums[2], xc=sums[3] + sums[4])
        return (sums[5], fv), self.model.unravel(tbar)
