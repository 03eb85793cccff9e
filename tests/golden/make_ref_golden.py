"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES (unmodified, imported from
/root/reference under the test-only jax/flax/distrax stand-ins of tests/_jaxstub; loader: tests/_refload.py).

    python tests/golden/make_ref_golden.py            # all cases (about ten minutes on 8 cores)
    python tests/golden/make_ref_golden.py tables functionals rhs step_h2    # a subset

What runs from the reference, as written: functionals.py (every factory of a11-a19, f1), utils.py (coordinates,
one_hot_encode, batch_generator, batche_generator_1D), promolecular_distrax.py, hgh_pseudopotentials.py (tables),
equiv_flows.py / cn_flows.py (the Flax modules), jax_ode.py (neural_ode, neural_ode_score with its nested-autodiff
`_evol_fn_i`).  What is restated because it is NOT in /root/reference: jax.experimental.ode.odeint, flax.linen, distrax
(tests/_jaxstub), and the `loss` closure of the drivers (a closure inside `training()`, so it cannot be imported; it is
re-typed below statement by statement with the driver line it follows).  The parameter gradient is reverse-mode
autodiff through the adaptive solve (the reference uses odeint's continuous adjoint; both are exact to O(tol)).

Inputs (theta, base samples) are drawn with NumPy generators and stored in every fixture: JAX's PRNG is not reproduced.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import _refload  # noqa: E402

ref = _refload.load_reference()
if ref is None:
    raise SystemExit("the reference checkout is not present; fixtures can only be generated where /root/reference exists")

import torch  # noqa: E402
import jax  # noqa: E402  (tests/_jaxstub)
import jax.numpy as jnp  # noqa: E402
import jax.random as jrnd  # noqa: E402
from jax import lax  # noqa: E402
from jax.flatten_util import ravel_pytree  # noqa: E402
from jax.experimental import ode as stub_ode  # noqa: E402

F = ref.functionals
BOHR = 1.8897259886
T64 = torch.float64


def npy(x):
    return x.detach().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    arrs = {k: npy(v) for k, v in kw.items()}
    for k, a in arrs.items():
        # parameters are float32-representable by construction; the 527 k-entry 1-D gradient is only compared at 1e-4
        if k.endswith("theta") or (k == "grad" and a.size > 100000):
            arrs[k] = a.astype(np.float32)
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}.npz ({os.path.getsize(path) / 1024:.0f} KB)", flush=True)


# ----------------------------------------------------------------------------------------------------------------
# molecule setups exactly as each driver builds them
# ----------------------------------------------------------------------------------------------------------------

def driver_setup(name):
    """-> (Ne, z, coords, z_one_hot) the way the named driver constructs them."""
    if name == "H2":      # H2_mol_ofdft_min_equiv_flows.py:108-118
        coords = jnp.array([[0., 0., -1.4008538753 / 2], [0., 0., 1.4008538753 / 2]]) * BOHR
        z = jnp.array([1., 1.])
        Ne = 2
    elif name == "LiH":   # LiH_mol_ofdft_min_equiv_flows.py:115-126
        coords = jnp.array([[0., 0., -1.5949 / 2], [0., 0., 1.5949 / 2]]) * BOHR
        z = jnp.array([3., 1.])
        Ne = 4
    elif name == "H2O":   # H20_mol_ofdft_min_equiv_flows.py:70-84
        coords = jnp.array([[0.0, 0.0, 0.1189120], [0.0, 0.7612710, -0.4756480], [0.0, -0.7612710, -0.4756480]]) * BOHR
        z = jnp.array([8., 1., 1.])
        Ne = 10
    else:                 # OFDFT_NF.py:66-73: coordinates(mol_name) + one_hot_encode
        Ne, atoms, z, coords = ref.utils.coordinates(name)
        return Ne, z, coords, ref.utils.one_hot_encode(z)
    n_atoms_type = jnp.unique(z).shape[0] + 1
    return Ne, z, coords, jax.nn.one_hot(z, n_atoms_type)


def make_theta(rng, n_in, features, n_out=1, last_scale=0.5, bias=0.1):
    """Flat parameters in ravel_pytree key order (last_layer.{bias,kernel}, layers_i.{bias,kernel}); non-zero last layer and
    biases so that every gradient entry is exercised (the reference's zero-init makes the flow the identity)."""
    parts = [np.full(n_out, 0.05 if last_scale > 0 else 0.0),
             last_scale * rng.standard_normal((features[-1], n_out)).ravel() / np.sqrt(features[-1])]
    prev = n_in
    for h in features:
        parts += [bias * rng.standard_normal(h), rng.standard_normal((prev, h)).ravel() / np.sqrt(prev)]
        prev = h
    return np.concatenate(parts).astype(np.float32).astype(np.float64)


def gnn_models(features, coords, onehot):
    GCNF = ref.equiv_flows.Gen_EqvFlow
    model_rev = GCNF(3, features, xyz_nuclei=coords, z_one_hot=onehot, bool_neg=False)
    model_fwd = GCNF(3, features, xyz_nuclei=coords, z_one_hot=onehot, bool_neg=True)
    test_inputs = lax.concatenate((jnp.ones((1, 3)), jnp.ones((1, 1))), 1)
    params = model_rev.init(jrnd.PRNGKey(0), jnp.array(0.), test_inputs)      # H20_...py:90-91
    return model_rev, model_fwd, params


def mlp_models(features, d=1):
    CNF = ref.cn_flows.Gen_CNFSimpleMLP
    model_rev = CNF(d, features, bool_neg=False)
    model_fwd = CNF(d, features, bool_neg=True)
    test_inputs = lax.concatenate((jnp.ones((1, d)), jnp.ones((1, 1))), 1)
    params = model_rev.init(jrnd.PRNGKey(0), jnp.array(0.), test_inputs)      # LiH.py:64-67
    return model_rev, model_fwd, params


def tree_paths(tree, prefix=""):
    out = []
    for k in sorted(tree):
        if isinstance(tree[k], dict):
            out += tree_paths(tree[k], prefix + "/" + k)
        else:
            out.append(f"{prefix}/{k}:{'x'.join(str(s) for s in tree[k].shape)}")
    return out


def capture_rhs(fn_ode, params, states, model, t_values, d):
    """The right-hand side the reference hands to odeint (jax_ode.py:36-37 `_evol_fun` / :95-98 `_evol_fn_i` vmapped) is a
    closure; catch it by swapping the module-level `odeint` name of jax_ode for a recorder."""
    jo = ref.jax_ode
    saved = jo.odeint
    box = {}

    def recorder(func, y0, ts, **kw):
        box["f"] = func
        return torch.stack([y0, y0], 0)
    jo.odeint = recorder
    try:
        fn_ode(params, states, model, 0., 1., d)
    finally:
        jo.odeint = saved
    return [box["f"](states, torch.as_tensor(t, dtype=T64)) for t in t_values]


# ----------------------------------------------------------------------------------------------------------------
# cases
# ----------------------------------------------------------------------------------------------------------------

def gen_tables():
    out = {}
    for m in ("H2", "LiH", "H2O", "CH4", "C6H6", "C27H46O"):
        Ne, atoms, z, coords = ref.utils.coordinates(m)          # utils.py:276-418
        out[f"{m}_Ne"] = Ne
        out[f"{m}_z"] = z
        out[f"{m}_coords"] = coords
        out[f"{m}_atoms"] = np.array(atoms)
        out[f"{m}_one_hot_encode"] = ref.utils.one_hot_encode(z)   # utils.py:251-274
    for m in ("H2", "LiH", "H2O"):
        Ne, z, coords, oh = driver_setup(m)
        out[f"{m}_driver_z"] = z
        out[f"{m}_driver_coords"] = coords
        out[f"{m}_driver_onehot"] = oh
    hp = ref.hgh_pseudopotentials
    keys = ("Zion", "rloc", "C1", "C2", "C3", "C4")
    out["hgh_H"] = np.array([hp.H_pp_params[k] for k in keys])     # hgh_pseudopotentials.py:8-9
    out["hgh_O"] = np.array([hp.O_pp_params[k] for k in keys])     # :10-11
    save("ref_tables", **out)


def gen_functionals():
    rng = np.random.default_rng(20240)
    n = 96
    # densities across the range the flows produce, including values at / below the 1e-30 clip of b88 / vwn / pw92
    logden = rng.uniform(-25.0, 2.0, n)
    logden[:4] = [-80.0, -69.5, -40.0, 1.5]
    den = np.exp(logden)[:, None]
    score = rng.standard_normal((n, 3)) * np.exp(rng.uniform(-3, 2, (n, 1)))
    score1 = score[:, :1].copy()
    x = rng.standard_normal((n, 3)) * 1.5
    xp = rng.standard_normal((n, 3)) * 1.5
    xp[:3] = x[:3] + np.array([[1e-3, 0, 0], [0, 0, 0], [1e-6, 1e-6, 0]])     # near-coincident pairs (eps matters)
    x1, xp1 = x[:, :1].copy(), xp[:, :1].copy()
    out = dict(den=den, score=score, score1=score1, x=x, xp=xp, x1=x1, xp1=xp1)
    tden, tscore, tscore1, tx, txp, tx1, txp1 = (torch.as_tensor(a) for a in (den, score, score1, x, xp, x1, xp1))

    def with_grads(tag, fn, args, diff):
        """value + d(sum)/d(arg) through the reference function (jax.grad of the stub = torch autograd)."""
        val = fn(*args)
        out[tag] = val
        for idx in diff:
            g = jax.grad(lambda a, idx=idx: jnp.sum(fn(*[a if i == idx else v for i, v in enumerate(args)])))(args[idx])
            out[f"{tag}__g{idx}"] = g

    for Ne in (2, 10, 42):
        for name in ("tf", "w", "tf-w"):
            with_grads(f"kin_{name}_Ne{Ne}", lambda d_, s_, Ne=Ne, name=name: F._kinetic(name)(d_, s_, Ne), (tden, tscore), (0, 1))
        for name in ("tf1d", "w1d", "tfw_1d"):
            with_grads(f"kin_{name}_Ne{Ne}", lambda d_, s_, Ne=Ne, name=name: F._kinetic(name)(d_, s_, Ne), (tden, tscore1), (0, 1))
        for name in ("lda", "b88_x_e", "dirac_b88_x_e"):
            with_grads(f"x_{name}_Ne{Ne}", lambda d_, s_, Ne=Ne, name=name: F._exchange_correlation(name)(d_, s_, Ne), (tden, tscore), (0, 1))
        for name in ("vwn_c_e", "pw92_c_e", "xc_1d"):
            with_grads(f"c_{name}_Ne{Ne}", lambda d_, Ne=Ne, name=name: F._exchange_correlation(name)(d_, Ne), (tden,), (0,))
        for name in ("hartree", "hartree_mt"):
            with_grads(f"hart_{name}_Ne{Ne}", lambda a, b, Ne=Ne, name=name: F._hartree(name)(a, b, Ne), (tx, txp), (0, 1))
        with_grads(f"hart_softc_Ne{Ne}", lambda a, b, Ne=Ne: F._hartree("softc")(a, b, Ne), (tx1, txp1), (0, 1))
    for m in ("H2", "LiH", "H2O", "C6H6"):
        Ne, z, coords, _ = driver_setup(m)
        mol = {"coords": coords, "z": z}                          # H20_...py:76
        for name in ("nuclei_potential", "hgh"):
            with_grads(f"nuc_{name}_{m}", lambda a, Ne=Ne, mol=mol, name=name: F._nuclear(name)(a, Ne, mol), (tx,), (0,))
    for (R, Za, Zb, Ne) in ((0.7, 3, 1, 2), (1.4, 1, 1, 2)):       # LiH.py:218-223 defaults + an H2-like pair
        with_grads(f"nuc_attr_R{R}_Za{Za}_Zb{Zb}", lambda a, R=R, Za=Za, Zb=Zb, Ne=Ne: F._nuclear("attr")(a, R, Za, Zb, Ne), (tx1,), (0,))
    save("ref_functionals", **out)


def gen_prior():
    out = {}
    for m in ("H2", "H2O", "C6H6"):
        Ne, z, coords, _ = driver_setup(m)
        prior = ref.promolecular_distrax.ProMolecularDensity(z.ravel(), coords)       # H20_...py:106
        batch = next(ref.utils.batch_generator(jrnd.PRNGKey(7), 64, prior))            # utils.py:76-108
        out[f"{m}_batch"] = batch
        out[f"{m}_prob"] = prior.prob(batch[:, :3])
    prior1 = sys.modules["distrax"].MultivariateNormalDiag(jnp.zeros(1), 1. * jnp.ones(1))   # LiH.py:78
    out["normal1d_batch"] = next(ref.utils.batche_generator_1D(jrnd.PRNGKey(8), 64, prior1))  # utils.py:110-143
    save("ref_prior", **out)


def gen_rhs():
    """f.apply(params, t, states) (equiv_flows.py:124-127 / cn_flows.py:179-182) and the score-augmented right-hand side
    (jax_ode.py:80-98), for every molecule setup and hidden-layer choice the drivers expose (OFDFT_NF.py:320-326)."""
    rng = np.random.default_rng(777)
    out = {}
    tvals = (0.0, 0.37, 1.0)
    out["t_values"] = np.array(tvals)
    for m, feats in (("H2", (64, 64)), ("LiH", (64, 64)), ("H2O", (64, 64)), ("C6H6", (64, 64)), ("CH4", (64, 64)),
                     ("H2O", (64,)), ("H2O", (64, 64, 64)), ("C6H6", (64,)), ("C6H6", (64, 64, 64)), ("C27H46O", (64, 64))):
        Ne, z, coords, oh = driver_setup(m)
        model_rev, model_fwd, params = gnn_models(feats, coords, oh)
        flat0, unravel = ravel_pytree(params)
        tag = f"{m}_L{len(feats)}"
        out[f"{tag}_tree"] = np.array(tree_paths(params))
        theta = make_theta(rng, 2 + oh.shape[1], feats)
        assert theta.size == flat0.numel(), (theta.size, flat0.numel())
        p = unravel(torch.as_tensor(theta))
        n = 24 if m != "C27H46O" else 12
        prior = ref.promolecular_distrax.ProMolecularDensity(z.ravel(), coords)
        states = next(ref.utils.batch_generator(jrnd.PRNGKey(11), n // 2, prior))     # (n, 7)
        out[f"{tag}_theta"], out[f"{tag}_states"] = theta, states
        out[f"{tag}_coords"], out[f"{tag}_onehot"] = coords, oh
        for neg, model in ((0, model_rev), (1, model_fwd)):
            out[f"{tag}_apply_neg{neg}"] = torch.stack([model.apply(p, torch.as_tensor(t, dtype=T64), states[:, :4]) for t in tvals])
            out[f"{tag}_rhs_score_neg{neg}"] = torch.stack(capture_rhs(ref.jax_ode.neural_ode_score, p, states, model, tvals, 3))
        print("rhs", tag, flush=True)
    for feats in ((512, 512, 512), (64, 64), (128, 128, 128, 128)):
        model_rev, model_fwd, params = mlp_models(feats)
        flat0, unravel = ravel_pytree(params)
        tag = f"mlp1d_{'x'.join(map(str, feats))}"
        out[f"{tag}_tree"] = np.array(tree_paths(params))
        theta = make_theta(rng, 2, feats)
        assert theta.size == flat0.numel()
        p = unravel(torch.as_tensor(theta))
        prior1 = sys.modules["distrax"].MultivariateNormalDiag(jnp.zeros(1), 1. * jnp.ones(1))
        states = next(ref.utils.batche_generator_1D(jrnd.PRNGKey(12), 12, prior1))     # (24, 3)
        out[f"{tag}_theta"], out[f"{tag}_states"] = theta, states
        for neg, model in ((0, model_rev), (1, model_fwd)):
            out[f"{tag}_apply_neg{neg}"] = torch.stack([model.apply(p, torch.as_tensor(t, dtype=T64), states[:, :2]) for t in tvals])
            out[f"{tag}_rhs_score_neg{neg}"] = torch.stack(capture_rhs(ref.jax_ode.neural_ode_score, p, states, model, tvals, 1))
        print("rhs", tag, flush=True)
    save("ref_rhs", **out)


def loss_3d(params, u_samples, model_fwd, batch_size, Ne, mol, names):
    """The `loss` closure of the 3-D drivers, statement by statement (H20_mol_ofdft_min_equiv_flows.py:127-130, 150-173)."""
    tw_kin, v_pot, h_pot, x_pot, c_pot = names
    t_functional = F._kinetic(tw_kin)                         # :139-143
    v_functional = F._nuclear(v_pot)
    vh_functional = F._hartree(h_pot)
    x_functional = F._exchange_correlation(x_pot)
    zt, logp_zt, score_zt = ref.jax_ode.neural_ode_score(params, u_samples, model_fwd, 0., 1., 3)   # :101-103
    den_all, x_all, score_all = jnp.exp(logp_zt), zt, score_zt                                        # :127-130
    den, denp = den_all[:batch_size], den_all[batch_size:]                                            # :154
    x, xp = x_all[:batch_size], x_all[batch_size:]
    score, scorep = score_all[:batch_size], score_all[batch_size:]
    e_t = t_functional(den, score, Ne)                                                                # :158
    e_h = vh_functional(x, xp, Ne)
    e_nuc_v = v_functional(x, Ne, mol)
    e_x = x_functional(den, score, Ne)
    if c_pot:
        e_c = F._exchange_correlation(c_pot)(den, Ne)                                                 # :162
    else:
        e_c = jnp.zeros_like(e_t)         # north-star configs 2-4 run without a correlation term
    e = e_t + e_nuc_v + e_h + e_x + e_c                                                               # :165
    energy = jnp.mean(e)
    f_values = torch.stack([jnp.mean(e_t), jnp.mean(e_nuc_v), jnp.mean(e_h), jnp.mean(e_x), jnp.mean(e_c), energy])
    return energy, (f_values, torch.cat((zt, logp_zt, score_zt), 1))


def loss_1d(params, u_samples, model_fwd, batch_size, Ne, R, Z_alpha, Z_beta, names):
    """LiH.py:96-99, 118-138."""
    tw_kin, v_pot, h_pot, xc_pot = names
    t_functional, v_functional = F._kinetic(tw_kin), F._nuclear(v_pot)
    vh_functional = F._hartree(h_pot)
    zt, logp_zt, score_zt = ref.jax_ode.neural_ode_score(params, u_samples, model_fwd, 0., 1., 1)
    den_all, x_all, score_all = jnp.exp(logp_zt), zt, score_zt
    den, x, xp, score = den_all[:batch_size], x_all[:batch_size], x_all[batch_size:], score_all[:batch_size]
    e_t = t_functional(den, score, Ne)
    e_h = vh_functional(x, xp, Ne)
    e_nuc_v = v_functional(x, R, Z_alpha, Z_beta, Ne)
    e_xc = F._exchange_correlation(xc_pot)(den, Ne) if xc_pot else jnp.zeros_like(e_t)
    e = e_t + e_h + e_nuc_v + e_xc
    energy = jnp.mean(e)
    f_values = torch.stack([jnp.mean(e_t), jnp.mean(e_nuc_v), jnp.mean(e_h), torch.zeros((), dtype=T64), jnp.mean(e_xc), energy])
    return energy, (f_values, torch.cat((zt, logp_zt, score_zt), 1))


def value_and_grad_flat(loss_fn, unravel, theta):
    """jax.value_and_grad(loss, has_aux=True)(params, batch) (H20_...py:177-178) on the flat parameter vector."""
    th = torch.as_tensor(theta, dtype=T64).clone().requires_grad_(True)
    energy, (terms, y1) = loss_fn(unravel(th))
    g, = torch.autograd.grad(energy, th)
    return float(energy), npy(terms), npy(g), npy(y1), dict(stub_ode.last_stats)


def step_case_3d(name, mol, feats, bs, names, seed, theta=None, batch=None, last_scale=0.5, rev_rows=0):
    t_start = time.time()
    rng = np.random.default_rng(seed)
    Ne, z, coords, oh = driver_setup(mol)
    model_rev, model_fwd, params = gnn_models(feats, coords, oh)
    _, unravel = ravel_pytree(params)
    if theta is None:
        theta = make_theta(rng, 2 + oh.shape[1], feats, last_scale=last_scale, bias=0.1 if last_scale > 0 else 0.0)
    if batch is None:
        prior = ref.promolecular_distrax.ProMolecularDensity(z.ravel(), coords)
        batch = npy(next(ref.utils.batch_generator(jrnd.PRNGKey(seed), bs, prior))).astype(np.float32).astype(np.float64)
    tb = torch.as_tensor(batch, dtype=T64)
    molinfo = {"coords": coords, "z": z}
    e, terms, g, y1, stats = value_and_grad_flat(lambda p: loss_3d(p, tb, model_fwd, bs, Ne, molinfo, names), unravel, theta)
    extra = {}
    if rev_rows:
        # rho_rev (H20_...py:132-137): reverse no-score flow of the transformed points back to the base density
        with torch.no_grad():
            p = unravel(torch.as_tensor(theta))
            xs = torch.as_tensor(y1[:rev_rows, :3])
            zt = lax.concatenate((xs, jnp.zeros((xs.shape[0], 1))), 1)
            z0, logp_z0 = ref.jax_ode.neural_ode(p, zt, model_rev, -1., 0., 3)
            prior = ref.promolecular_distrax.ProMolecularDensity(z.ravel(), coords)
            extra = dict(rev_x=xs, rev_z0=z0, rev_dlogp=logp_z0, rev_rho=jnp.exp(prior.log_prob(z0) - logp_z0))
    save(name, theta=theta.astype(np.float32), batch=batch.astype(np.float32), coords=coords, z=z, onehot=oh, Ne=Ne,
         features=np.array(feats), names=np.array([n or "" for n in names]), y1=y1, energy=e, terms=terms, grad=g,
         dopri_accepted=stats["accepted"], dopri_rejected=stats["rejected"], dopri_nfe=stats["nfe"], **extra)
    print(f"  {name}: E={e:.9f} terms={terms} {stats} ({time.time() - t_start:.0f} s)", flush=True)


def step_case_1d(name, feats, bs, names, seed, theta=None, batch=None):
    t_start = time.time()
    rng = np.random.default_rng(seed)
    model_rev, model_fwd, params = mlp_models(feats)
    _, unravel = ravel_pytree(params)
    if theta is None:
        theta = make_theta(rng, 2, feats)
    if batch is None:
        prior1 = sys.modules["distrax"].MultivariateNormalDiag(jnp.zeros(1), 1. * jnp.ones(1))
        batch = npy(next(ref.utils.batche_generator_1D(jrnd.PRNGKey(seed), bs, prior1))).astype(np.float32).astype(np.float64)
    tb = torch.as_tensor(batch, dtype=T64)
    Ne, R, Za, Zb = 2, 0.7, 3, 1                                  # LiH.py:211-223
    e, terms, g, y1, stats = value_and_grad_flat(lambda p: loss_1d(p, tb, model_fwd, bs, Ne, R, Za, Zb, names), unravel, theta)
    save(name, theta=theta.astype(np.float32), batch=batch.astype(np.float32), Ne=Ne, R=R, Z_alpha=Za, Z_beta=Zb,
         features=np.array(feats), names=np.array([n or "" for n in names]), y1=y1, energy=e, terms=terms, grad=g,
         dopri_accepted=stats["accepted"], dopri_rejected=stats["rejected"], dopri_nfe=stats["nfe"])
    print(f"  {name}: E={e:.9f} terms={terms} {stats} ({time.time() - t_start:.0f} s)", flush=True)


def gen_step_bench():
    """The benchmark's own parameters and base samples (bench.make_problem('h2o', 512)) through the reference: pins what
    the RK4 x16 fp32 kernels are timed on against the adaptive float64 solve."""
    import bench
    mol, nuc, theta, batch = bench.make_problem("h2o", 512)
    step_case_3d("ref_step_h2o_bench_bs512", "H2O", (64, 64), 512, ("tf-w", nuc, "hartree", "lda", ""), 0,
                 theta=theta.astype(np.float64), batch=batch.astype(np.float64))


CASES = {
    "tables": gen_tables,
    "functionals": gen_functionals,
    "prior": gen_prior,
    "rhs": gen_rhs,
    # north-star configs 2-4 (TF+W / nuclear / paired Hartree / LDA)
    "step_h2": lambda: step_case_3d("ref_step_h2_bs40", "H2", (64, 64), 40, ("tf-w", "nuclei_potential", "hartree", "lda", ""), 101),
    "step_lih": lambda: step_case_3d("ref_step_lih_bs32", "LiH", (64, 64), 32, ("tf-w", "nuclei_potential", "hartree", "lda", ""), 102),
    # the H2O driver's own defaults (H20_...py:234-244): tf-w / nuclei_potential|hgh / hartree / dirac_b88_x_e / vwn_c_e
    "step_h2o": lambda: step_case_3d("ref_step_h2o_hgh_b88_vwn_bs48", "H2O", (64, 64), 48, ("tf-w", "hgh", "hartree", "dirac_b88_x_e", "vwn_c_e"), 103, rev_rows=16),
    # OFDFT_NF.py defaults (:286-297): nuclei_potential / hartree / dirac_b88_x_e / pw92_c_e on benzene
    "step_c6h6": lambda: step_case_3d("ref_step_c6h6_b88_pw92_bs24", "C6H6", (64, 64), 24, ("tf-w", "nuclei_potential", "hartree", "dirac_b88_x_e", "pw92_c_e"), 104),
    # --nn 1 / --nn 3 (OFDFT_NF.py:320-326)
    "step_h2o_L1": lambda: step_case_3d("ref_step_h2o_L1_bs32", "H2O", (64,), 32, ("tf-w", "nuclei_potential", "hartree_mt", "lda", "pw92_c_e"), 105),
    "step_h2o_L3": lambda: step_case_3d("ref_step_h2o_L3_bs32", "H2O", (64, 64, 64), 32, ("tf-w", "nuclei_potential", "hartree", "b88_x_e", "vwn_c_e"), 106),
    # default initialisation: zero last layer => identity flow
    "step_h2o_zero": lambda: step_case_3d("ref_step_h2o_zero_init_bs32", "H2O", (64, 64), 32, ("tf-w", "hgh", "hartree", "lda", ""), 107, last_scale=0.0),
    # LiH.py defaults: tfw_1d / attr / softc / xc_1d on the 3 x 512 MLP flow
    "step_lih1d": lambda: step_case_1d("ref_step_lih1d_512x3_bs24", (512, 512, 512), 24, ("tfw_1d", "attr", "softc", "xc_1d"), 108),
    "step_mlp1d_small": lambda: step_case_1d("ref_step_mlp1d_64x2_bs40", (64, 64), 40, ("tfw_1d", "attr", "softc", ""), 109),
    "step_bench": gen_step_bench,
}


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    todo = sys.argv[1:] or list(CASES)
    for k in todo:
        print("==", k, flush=True)
        CASES[k]()
