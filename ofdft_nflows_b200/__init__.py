"""ofdft_nflows_b200 -- B200-native (sm_100a) Monte-Carlo energy-estimator hot path of ChemAI-Lab/ofdft_nflows.

Mirrors the reference package facade (ofdft_normflows/__init__.py:1-6) for the accelerated path:
``_kinetic, _nuclear, _hartree, _exchange_correlation, neural_ode, neural_ode_score, GCNF`` plus the
fused ``EnergyEstimator`` step.  Importing the package does not touch CUDA; the first compute call loads
``libofdft_b200.so`` and raises if it is missing (no CPU fallback).
"""
__version__ = "0.1.0"
