"""Host mirror of ofdft_normflows/equiv_flows.py: ``Gen_EqvFlow`` (the drivers import it as ``GCNF``).

Same constructor arguments, ``init`` / ``apply`` methods and parameter-pytree structure as the Flax module
(equiv_flows.py:108-127; pytree per SURVEY.md section 8b), but the arithmetic runs in the CUDA kernels of
libofdft_b200.so.  Parameters are torch CUDA tensors (device memory only; no torch arithmetic on the path
besides flattening the pytree).
"""
from __future__ import annotations

import math
from typing import Any, Dict, Sequence

import numpy as np
import torch

from . import ops

_LAYER_KEYS = ("bias", "kernel")


def _ravel(layers: Dict[str, Dict[str, torch.Tensor]]) -> torch.Tensor:
    """Key-sorted flattening (``jax.flatten_util.ravel_pytree`` order): last_layer.{bias,kernel}, layers_i.{bias,kernel}."""
    parts = []
    for name in sorted(layers.keys()):
        for leaf in _LAYER_KEYS:
            parts.append(layers[name][leaf].reshape(-1))
    return torch.cat(parts)


def _unravel(flat: torch.Tensor, n_in: int, features: Sequence[int], n_out: int) -> Dict[str, Dict[str, torch.Tensor]]:
    shapes = {"last_layer": {"bias": (n_out,), "kernel": (features[-1], n_out)}}
    prev = n_in
    for i, h in enumerate(features):
        shapes[f"layers_{i}"] = {"bias": (h,), "kernel": (prev, h)}
        prev = h
    out: Dict[str, Dict[str, torch.Tensor]] = {}
    o = 0
    for name in sorted(shapes.keys()):
        out[name] = {}
        for leaf in _LAYER_KEYS:
            shp = shapes[name][leaf]
            n = int(np.prod(shp))
            out[name][leaf] = flat[o:o + n].reshape(shp)
            o += n
    assert o == flat.numel(), (o, flat.numel())
    return out


def _init_dense_stack(seed: int, n_in: int, features: Sequence[int], n_out: int, device) -> Dict[str, Dict[str, torch.Tensor]]:
    """flax.linen.Dense defaults: LeCun-normal kernel, zero bias; zero-initialised last layer
    (equiv_flows.py:41-45, cn_flows.py:122-126) => the initial flow is the identity."""
    rng = np.random.default_rng(seed)
    layers = {}
    prev = n_in
    for i, h in enumerate(features):
        # truncated normal (+-2 sigma) rescaled to unit variance, as jax.nn.initializers.lecun_normal
        w = rng.standard_normal((prev, h))
        bad = np.abs(w) > 2.0
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2.0
        w = w / 0.87962566103423978 / math.sqrt(prev)
        layers[f"layers_{i}"] = {"kernel": torch.tensor(w, dtype=torch.float32, device=device),
                                 "bias": torch.zeros(h, dtype=torch.float32, device=device)}
        prev = h
    layers["last_layer"] = {"kernel": torch.zeros((prev, n_out), dtype=torch.float32, device=device),
                            "bias": torch.zeros(n_out, dtype=torch.float32, device=device)}
    return layers


class Gen_EqvFlow:
    """``Gen_EqvFlow(in_out_dim, features, xyz_nuclei, z_one_hot, bool_neg=False)`` (equiv_flows.py:108-114).

    g(x,t) = sum_i f_theta(y0 t, |x-R_i|, onehot_i) (x-R_i), exact trace, output scaled by y0 (-1 if bool_neg).
    ``in_out_dim == 3`` and ``features`` = one to three hidden layers of 64 units are instantiated in CUDA ((64, 64), what every
    driver but OFDFT_NF.py hard-codes, e.g. H20_mol_ofdft_min_equiv_flows.py:263, runs on the tensor-core kernels); other
    shapes raise NotImplementedError.
    """

    def __init__(self, in_out_dim: Any, features: Sequence[int], xyz_nuclei: Any, z_one_hot: Any, bool_neg: bool = False,
                 device: str = "cuda"):
        if int(in_out_dim) != 3:
            raise NotImplementedError("Gen_EqvFlow: only in_out_dim == 3 is supported")
        feats = tuple(int(f) for f in features)
        if not (1 <= len(feats) <= 3) or any(f != 64 for f in feats):
            # the drivers expose (64,), (64, 64) and (64, 64, 64) (OFDFT_NF.py:320-326); everything else hard-codes (64, 64)
            raise NotImplementedError("Gen_EqvFlow: CUDA kernels are instantiated for 1-3 hidden layers of 64 units")
        self.in_out_dim = 3
        self.features = feats
        self.bool_neg = bool(bool_neg)
        self.y0 = -1.0 if self.bool_neg else 1.0
        self.device = torch.device(device)
        self.xyz_nuclei = torch.as_tensor(np.asarray(xyz_nuclei, dtype=np.float32)).reshape(-1, 3).to(self.device).contiguous()
        oh = np.asarray(z_one_hot, dtype=np.float32)
        self.z_one_hot = torch.as_tensor(oh).reshape(self.xyz_nuclei.shape[0], -1).to(self.device).contiguous()
        self.n_types = self.z_one_hot.shape[1]
        self.n_in = 2 + self.n_types
        if self.xyz_nuclei.shape[0] > 128 or self.n_types > 8 or len({tuple(r) for r in oh.reshape(len(oh), -1).tolist()}) > 8:
            raise NotImplementedError("Gen_EqvFlow: at most 128 nuclei, 8 species and 8 distinct one-hot rows")

    # ---- Flax-like API -------------------------------------------------------------------------
    def init(self, key: Any = 0, t: Any = None, test_inputs: Any = None) -> Dict[str, Any]:
        """``model.init(key, jnp.array(0.), test_inputs)`` (H20_...py:90-91); ``key`` is an integer seed here."""
        layers = _init_dense_stack(int(key), self.n_in, self.features, 1, self.device)
        return {"params": {"cnf": {"net": {"VmapNN_0": layers}}}}

    def apply(self, params: Dict[str, Any], t: Any, states: torch.Tensor) -> torch.Tensor:
        """``f.apply(params, t, states)`` (jax_ode.py:37): d/dt [x, logp] for states (B, 4)."""
        jets = ops.JETS_XLOGP if states.shape[1] == 4 else ops.JETS_FULL
        return ops.gnn_rhs(self.flat_theta(params), self.xyz_nuclei, self.z_one_hot, states, jets, float(t), self.y0)

    # ---- pytree <-> flat theta -----------------------------------------------------------------
    @staticmethod
    def _layers(params: Dict[str, Any]) -> Dict[str, Dict[str, torch.Tensor]]:
        return params["params"]["cnf"]["net"]["VmapNN_0"]

    def flat_theta(self, params: Dict[str, Any]) -> torch.Tensor:
        return _ravel(self._layers(params))

    def unravel(self, flat: torch.Tensor) -> Dict[str, Any]:
        return {"params": {"cnf": {"net": {"VmapNN_0": _unravel(flat, self.n_in, self.features, 1)}}}}


GCNF = Gen_EqvFlow   # the drivers' alias (H20_mol_ofdft_min_equiv_flows.py: ``from ...equiv_flows import Gen_EqvFlow as GCNF``)
