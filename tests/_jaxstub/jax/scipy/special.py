"""Only imported by hgh_pseudopotentials.py:4 for its unfinished projector stubs (out of scope); the local-term tables
(:8-11) are what the hot path reads."""


def sph_harm(*a, **k):
    raise NotImplementedError("jax.scipy.special.sph_harm is outside the hot path")


def gamma(*a, **k):
    raise NotImplementedError("jax.scipy.special.gamma is outside the hot path")
