from . import prng  # noqa: F401
