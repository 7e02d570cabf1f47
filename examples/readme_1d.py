"""The README example of alonfnt/bayex (README.md:32-55) with the import line changed - BASELINE config 1.

    python examples/readme_1d.py

f(x) = -(1.4 - 3x) sin(18x) on Real(0, 2), three initial points, 20 sample/fit steps with the PI acquisition.
Needs a B200 (there is no CPU fallback)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bayex_b200 as bayex


def f(x):
    return -(1.4 - 3 * x) * np.sin(18 * x)


domain = {'x': bayex.domain.Real(0.0, 2.0)}
optimizer = bayex.Optimizer(domain=domain, maximize=True, acq='PI')

# Define some prior evaluations to initialise the GP.
params = {'x': [0.0, 0.5, 1.0]}
ys = [f(x) for x in params['x']]
opt_state = optimizer.init(ys, params)

# Sample new points using a Threefry key (bayex.random.key mirrors jax.random.key) and the optimizer state.
ori_key = bayex.random.key(42)
for step in range(20):
    key = bayex.random.fold_in(ori_key, step)
    new_params = optimizer.sample(key, opt_state)
    y_new = f(**new_params)
    opt_state = optimizer.fit(opt_state, y_new, new_params)
    print(f"step {step:2d}: x = {float(new_params['x'][0]):.4f}  f(x) = {float(np.ravel(y_new)[0]):+.4f}  best so far {opt_state.best_score:+.4f}")

print("best:", {k: float(v) for k, v in opt_state.best_params.items()}, "score", opt_state.best_score)
