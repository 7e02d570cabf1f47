"""Flat on-disk format for ``OptimizerState`` (the reference has none - optimizer.py:14-36).

One ``.npz`` with the padded buffers exactly as the state holds them (dtypes included: int32 Integer storage before the
first ``fit``, float32 after - quirk Q14), so ``load_state(save_state(s))`` is field-for-field identical and the
optimiser's posterior cache keys (a digest of these arrays) still match.  The GPU-resident factor is *not* stored: it
is rebuilt from the state on first use (``Optimizer.posterior``), deterministically.
"""
from __future__ import annotations

import numpy as np

from .gp import GPParams
from .optimizer import OptimizerState

FORMAT_VERSION = 1


def save_state(path, state: OptimizerState) -> None:
    names = list(state.params.keys())
    arrays = {"__version__": np.int32(FORMAT_VERSION), "__names__": np.array(names, dtype=np.str_),
              "ys": np.asarray(state.ys), "mask": np.asarray(state.mask), "best_score": np.float32(state.best_score),
              "gp_noise": np.asarray(state.gp_params.noise, dtype=np.float32),
              "gp_amplitude": np.asarray(state.gp_params.amplitude, dtype=np.float32),
              "gp_lengthscale": np.asarray(state.gp_params.lengthscale, dtype=np.float32)}
    for i, k in enumerate(names):
        arrays[f"param_{i}"] = np.asarray(state.params[k])
        arrays[f"best_{i}"] = np.asarray(state.best_params[k])
    with open(path, "wb") as f:
        np.savez(f, **arrays)


def load_state(path) -> OptimizerState:
    with np.load(path, allow_pickle=False) as z:
        version = int(z["__version__"])
        if version != FORMAT_VERSION:
            raise ValueError(f"unsupported OptimizerState file version {version} (this build reads {FORMAT_VERSION})")
        names = [str(s) for s in z["__names__"]]
        params = {k: z[f"param_{i}"] for i, k in enumerate(names)}
        best = {k: z[f"best_{i}"][()] for i, k in enumerate(names)}
        gp = GPParams(z["gp_noise"], z["gp_amplitude"], z["gp_lengthscale"])
        return OptimizerState(params=params, ys=z["ys"], best_score=float(z["best_score"]), best_params=best,
                              mask=z["mask"], gp_params=gp)
