from . import ode  # noqa: F401
