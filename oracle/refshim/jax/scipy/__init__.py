from . import linalg, special, stats  # noqa: F401
