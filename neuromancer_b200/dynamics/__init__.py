from . import integrators, ode  # noqa: F401
