"""ofdft_nflows_b200 -- B200-native (sm_100a) Monte-Carlo energy-estimator hot path of ChemAI-Lab/ofdft_nflows.

Mirrors the reference package facade (ofdft_normflows/__init__.py:1-6) for the accelerated path:
``_kinetic, _nuclear, _hartree, _exchange_correlation, neural_ode, neural_ode_score, GCNF`` plus the
fused ``EnergyEstimator`` step.  Importing the package does not touch CUDA; the first compute call loads
``libofdft_b200.so`` and raises if it is missing (no CPU fallback).
"""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy re-exports of the reference facade (ofdft_normflows/__init__.py:1-6); nothing here touches CUDA at import time
    if name in ("_kinetic", "_nuclear", "_hartree", "_exchange_correlation"):
        from . import functionals
        return getattr(functionals, name)
    if name in ("neural_ode", "neural_ode_score"):
        from . import jax_ode
        return getattr(jax_ode, name)
    if name in ("GCNF", "Gen_EqvFlow"):
        from . import equiv_flows
        return getattr(equiv_flows, name)
    if name in ("CNF", "Gen_CNFSimpleMLP"):
        from . import cn_flows
        return getattr(cn_flows, name)
    if name in ("ProMolecularDensity", "batch_generator", "batche_generator_1D", "coordinates", "one_hot_encode"):
        from . import utils
        return getattr(utils, name)
    if name in ("EnergyEstimator", "F_values"):
        from . import estimator
        return getattr(estimator, name)
    raise AttributeError(name)
