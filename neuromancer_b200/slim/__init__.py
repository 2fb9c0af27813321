from .linear import Linear, LinearBase, maps  # noqa: F401
