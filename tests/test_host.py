"""CPU tests of the host-side mirror of the reference interface (no CUDA needed)."""
import numpy as np
import pytest
import torch

from oracle import ofdft_oracle as O
from ofdft_nflows_b200 import ops, utils
from ofdft_nflows_b200.equiv_flows import Gen_EqvFlow, GCNF, _ravel, _unravel


def test_pytree_matches_flax_structure_and_ravel_order():
    Ne, atoms, z, coords = utils.coordinates("H2O")
    m = GCNF(3, (64, 64), coords, utils.jax_one_hot(z), bool_neg=True, device="cpu")
    params = m.init(0)
    layers = params["params"]["cnf"]["net"]["VmapNN_0"]
    assert sorted(layers) == ["last_layer", "layers_0", "layers_1"]
    assert layers["layers_0"]["kernel"].shape == (5, 64) and layers["layers_1"]["kernel"].shape == (64, 64)
    assert layers["last_layer"]["kernel"].shape == (64, 1)
    assert float(layers["last_layer"]["kernel"].abs().sum()) == 0.0          # zero-init last layer (equiv_flows.py:43-45)
    assert float(layers["layers_0"]["bias"].abs().sum()) == 0.0
    flat = m.flat_theta(params)
    assert flat.numel() == ops.gnn_theta_size(3) == 4609
    # key-sorted order == the oracle's flat layout
    p = O.unflatten_params(flat.numpy().astype(np.float64), 5, (64, 64))
    assert np.array_equal(p.hidden[0][0], layers["layers_0"]["kernel"].numpy().astype(np.float64))
    assert np.array_equal(p.hidden[1][0], layers["layers_1"]["kernel"].numpy().astype(np.float64))
    back = m.unravel(flat)
    assert torch.equal(back["params"]["cnf"]["net"]["VmapNN_0"]["layers_1"]["kernel"], layers["layers_1"]["kernel"])
    # LeCun-normal scale
    assert abs(float(layers["layers_1"]["kernel"].std()) - 1 / 8) < 0.01


def test_unsupported_shapes_raise():
    Ne, atoms, z, coords = utils.coordinates("H2")
    with pytest.raises(NotImplementedError):
        Gen_EqvFlow(3, (512, 512), coords, utils.jax_one_hot(z), device="cpu")
    with pytest.raises(NotImplementedError):
        Gen_EqvFlow(2, (64, 64), coords, utils.jax_one_hot(z), device="cpu")


def test_one_hot_rules_and_geometry_match_oracle():
    for name in ("H2", "LiH", "H2O", "C6H6"):
        Ne, atoms, z, coords = utils.coordinates(name)
        mol = O.molecule(name)
        assert Ne == mol["Ne"] and np.allclose(coords, mol["coords"]) and np.array_equal(z, mol["z"])
        oh = utils.one_hot_encode(z) if name == "C6H6" else utils.jax_one_hot(z)
        assert np.array_equal(oh, mol["onehot"])
    assert np.array_equal(utils.jax_one_hot(np.array([8., 1., 1.])), np.array([[0, 0, 0], [0, 1, 0], [0, 1, 0]], dtype=np.float32))
    assert np.array_equal(utils.jax_one_hot(np.array([3., 1.]))[0], np.zeros(3, dtype=np.float32))


def test_batch_generator_layout_and_prior():
    Ne, atoms, z, coords = utils.coordinates("H2O")
    prior = utils.ProMolecularDensity(z, coords)
    b = next(utils.batch_generator(0, 64, prior))
    assert b.shape == (128, 7) and b.dtype == np.float32
    lp, sc = O.promolecular_logp_score(b[:, :3].astype(np.float64), coords, z)
    assert np.abs(lp[:, 0] - b[:, 3]).max() < 1e-5 and np.abs(sc - b[:, 4:]).max() < 1e-5
    b1 = next(utils.batche_generator_1D(0, 16))
    assert b1.shape == (32, 3) and np.allclose(b1[:, 2], -b1[:, 0])


def test_energy_spec_mapping():
    Ne, atoms, z, coords = utils.coordinates("H2O")
    cs, c, zz = ops.EnergySpec(Ne=10, d=3, kin="TF-W", hart="hartree", nuc="HGH", x="lda", coords=coords, z=z).to_c()
    assert (cs.d, cs.kin, cs.hart, cs.nuc, cs.x, cs.c, cs.n_nuclei) == (3, 3, 1, 2, 1, 0, 3) and c.dtype == np.float32
    cs2, _, _ = ops.EnergySpec(Ne=10, d=3, x="dirac_b88_x_e", c="vwn_c_e", coords=coords, z=z).to_c()
    assert (cs2.x, cs2.c) == (3, 1)
    cs1, c1, z1 = ops.EnergySpec(Ne=2, d=1, kin="tfw_1d", hart="softc", nuc="attraction", x="", R=0.7, Z_alpha=3, Z_beta=1).to_c()
    assert (cs1.kin, cs1.hart, cs1.nuc, cs1.x) == (5, 2, 3, 0) and c1 is None
    with pytest.raises(NotImplementedError):
        ops.EnergySpec(Ne=2, kin="kinetic").to_c()


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops._f32c(torch.zeros(4))
