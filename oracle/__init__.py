"""CPU oracle for the bayex GP-posterior + acquisition hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker (or as
the timed CPU baseline), never as the thing shipped.  ``bayex_b200`` never
imports this package and fails loudly when its CUDA library is missing.

What it is: a NumPy/SciPy restatement of the reference's four modules
(``/root/reference/bayex/{gp,acq,domain,optimizer}.py``), each function citing
the reference file:line it follows, plus a restatement of the parts of JAX and
optax the reference leans on (Threefry-2x32 PRNG, ``uniform``/``randint``,
``ndtr``, ``clip_by_global_norm`` + ``adamw``).

Parity pinning status
---------------------
* GP / acquisition / NLML / state-machine arithmetic: PINNED.  JAX itself is
  not installable here, so ``oracle/make_golden.py`` executes the reference's
  *own source files* from ``/root/reference`` under a NumPy stand-in for the
  ``jax``/``optax`` namespaces (``oracle/_shim``) and stores the outputs in
  ``tests/golden/``; ``tests/test_oracle_golden.py`` holds this restatement to
  them, to the three closed-form ``exp_quadratic`` KATs of the reference's
  ``tests/test_gp.py:10-12`` and to the independent fp64 values of SURVEY §4.
* PRNG bit streams (Threefry, ``split``, ``fold_in``, ``uniform``, ``randint``):
  pinned to the Random123 known-answer vectors and to the ``split(key(0))``
  example of the JAX documentation; the reference's own tests pin ranges and
  dtypes only (``tests/test_domain.py:15-42``) => bit-level parity with a live
  JAX is UNPINNED beyond those vectors (jax is an unpinned dependency,
  ``pyproject.toml:20``).
* AdamW / global-norm clip (optax, unpinned dependency): restated from the
  published algorithm; PARITY UNPINNED (the reference asserts no fitted value).
* L-BFGS refinement (``optimizer.py:39-73``): PARITY UNPINNED and not restated;
  the contract tested instead is value-based (see ``oracle/optimizer.py``).
"""

from . import prng, gp, acq, domain, optimizer  # noqa: F401
