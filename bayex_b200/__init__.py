"""bayex_b200 - B200-native GP-posterior + acquisition hot path behind the bayex API.

``import bayex_b200 as bayex`` gives the reference's surface (bayex/__init__.py:1-11): ``Optimizer``,
``domain``, ``acq`` (plus ``gp`` and a JAX-free ``random`` for keys).
"""
from . import acq, domain, gp, random  # noqa: F401
from .optimizer import Optimizer, OptimizerState  # noqa: F401
from .state_io import load_state, save_state  # noqa: F401

__version__ = "0.2.2+b200.1"

__all__ = [
    "Optimizer",
    "domain",
    "acq",
]
