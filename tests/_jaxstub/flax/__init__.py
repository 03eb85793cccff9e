"""Test-only stand-in for the `flax` names the reference flow modules use (see ../README.md)."""
from . import linen  # noqa: F401
