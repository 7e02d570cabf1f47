"""``jax.scipy`` stand-in. TEST INFRASTRUCTURE ONLY."""
from . import linalg, stats  # noqa: F401
