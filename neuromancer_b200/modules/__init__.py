from . import activations, blocks, functions  # noqa: F401
