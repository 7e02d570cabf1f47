"""Host-side mirror of the reference interface (no GPU): keys, domains, state machine helpers, error behaviour."""
import numpy as np
import pytest

import bayex_b200 as bayex
from bayex_b200 import optimizer as bopt, random as brandom
from bayex_b200.domain import Integer, Real
from oracle import optimizer as oopt, prng


def test_key_helpers_match_oracle_threefry():
    assert brandom.key(42).tolist() == [0, 42] == prng.key(42).tolist()
    k = brandom.key(42)
    for step in range(5):
        k = brandom.fold_in(k, step)
        assert k.tolist() == prng.fold_in(prng.key(42) if step == 0 else kk, step).tolist()
        kk = k
    assert np.array_equal(brandom.split(brandom.key(7), 5), prng.split(prng.key(7), 5))
    assert brandom.as_key(3).tolist() == [0, 3]
    with pytest.raises(ValueError):
        brandom.as_key([1, 2, 3])


# -- reference tests/test_domain.py restated (the transform / eq / hash / assert parts need no device) --------------
@pytest.mark.parametrize("lower, upper", [(0.0, 1.0), (-5.5, 5.5)])
def test_real_transform_clips_within_bounds(lower, upper):
    out = Real(lower, upper).transform(np.array([lower - 1, (lower + upper) / 2, upper + 1]))
    assert np.all(out >= lower) and np.all(out <= upper)


@pytest.mark.parametrize("lower, upper", [(0, 5), (-3, 3)])
def test_integer_transform_and_type(lower, upper):
    out = Integer(lower, upper).transform(np.array([lower - 2.3, (lower + upper) / 2.0, upper + 2.7]))
    assert np.issubdtype(out.dtype, np.floating) and out.dtype == np.float32
    assert np.all(out >= lower) and np.all(out <= upper) and np.all(out == np.round(out))
    assert Integer(-3, 3).transform(np.array([0.5, 1.5, 2.5, -0.5])).tolist() == [0.0, 2.0, 2.0, -0.0]  # half-to-even


def test_domain_equality_and_hash():
    a, b, c = Real(0.0, 1.0), Real(0.0, 1.0), Real(1.0, 2.0)
    assert a == b and a != c and hash(a) == hash(b) and hash(a) != hash(c)
    i1, i2, i3 = Integer(1, 5), Integer(1, 5), Integer(0, 4)
    assert i1 == i2 and i1 != i3 and hash(i1) == hash(i2) and hash(i1) != hash(i3)
    assert Real(0, 1).dtype == np.float32 and Integer(0, 1).dtype == np.int32


@pytest.mark.parametrize("lower, upper", [(1.0, 1.0), ("a", 1.0), (1.0, "b"), (5.0, 3.0)])
def test_real_init_invalid(lower, upper):
    with pytest.raises((AssertionError, TypeError)):
        Real(lower, upper)


@pytest.mark.parametrize("lower, upper", [(5, 5), ("a", 5), (0, "b"), (10, 1)])
def test_integer_init_invalid(lower, upper):
    with pytest.raises(AssertionError):
        Integer(lower, upper)


def test_unknown_acq_raises_reference_message():
    for acq in ["EI", "PI", "UCB", "LCB"]:
        bayex.Optimizer(domain={"x": bayex.domain.Real(-2, 2)}, acq=acq)
    for acq in ["random", "magic"]:
        with pytest.raises(ValueError, match=f"Acquisition function {acq} is not implemented"):
            bayex.Optimizer(domain={"x": bayex.domain.Real(-2, 2)}, acq=acq)


def test_public_surface_matches_reference_exports():
    assert bayex.__all__ == ["Optimizer", "domain", "acq"]
    assert bayex.optimizer.OptimizerState._fields == ("params", "ys", "best_score", "best_params", "mask", "gp_params")
    assert bayex.gp.GPParams._fields == ("noise", "amplitude", "lengthscale")
    for name in ("expected_improvement", "probability_improvement", "upper_confidence_bounds", "lower_confidence_bounds"):
        assert callable(getattr(bayex.acq, name))
    for name in ("MASK_VARIANCE", "GPState", "exp_quadratic", "cov", "softplus", "gaussian_process",
                 "marginal_likelihood", "predict", "grad_fun", "neg_log_likelihood", "posterior_fit"):
        assert hasattr(bayex.gp, name)


def test_expand_doubles_buffers_like_reference():
    opt = bayex.Optimizer({"x": Real(0, 1), "k": Integer(0, 9)})
    oo = oopt.Optimizer({"x": __import__("oracle").domain.Real(0, 1), "k": __import__("oracle").domain.Integer(0, 9)})
    st = bopt.OptimizerState({"x": np.arange(10, dtype=np.float32), "k": np.arange(10, dtype=np.int32)},
                             np.arange(10, dtype=np.float32), 9.0, {}, np.ones(10, bool), None)
    a, b = opt.expand(st), oo.expand(oopt.OptimizerState(*st))
    assert a.mask.shape == b.mask.shape == (20,) and np.array_equal(a.mask, b.mask)
    assert np.array_equal(a.ys, b.ys) and a.params["k"].dtype == np.int32
    st2 = st._replace(mask=np.array([True] * 9 + [False]))
    assert opt.expand(st2).mask.shape == (10,)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    opt = bayex.Optimizer({"x": Real(0, 1)})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.init([0.1, 0.2], {"x": [0.0, 1.0]})


def test_shard_range_partitions_the_candidate_axis():
    for size in (1, 7, 10_000, 2 ** 24):
        for world in (1, 2, 3, 4, 8):
            spans = [bopt.shard_range(size, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == size
            pos = 0
            for b, c in spans:
                assert b == pos or c == 0
                pos += c


def test_merge_topk_equals_global_stable_argsort():
    rng = np.random.default_rng(0)
    vals = rng.standard_normal(4000).astype(np.float32)
    vals[rng.integers(0, 4000, 300)] = 0.25  # ties
    vals[[5, 77]] = np.nan
    vals[100] = -0.0
    want = oopt.jnp_argsort(vals)[-50:]
    parts_v, parts_i = [], []
    for r in range(4):
        b, c = bopt.shard_range(4000, r, 4)
        loc = oopt.jnp_argsort(vals[b:b + c])[-50:] + b
        parts_v.append(vals[loc]); parts_i.append(loc.astype(np.int32))
    v, i = bopt.merge_topk(np.concatenate(parts_v), np.concatenate(parts_i), 50)
    assert np.array_equal(i, want)


def test_state_save_load_roundtrip(tmp_path):
    """flat .npz format; dtypes (int32 Integer storage before the first fit) survive."""
    from bayex_b200.gp import GPParams
    st = bayex.OptimizerState(
        params={"a": np.array([0.5, 1.5, 0, 0], np.float32), "k": np.array([3, -2, 0, 0], np.int32)},
        ys=np.array([1.0, -2.0, 0, 0], np.float32), best_score=1.0, best_params={"a": np.float32(0.5), "k": np.int32(3)},
        mask=np.array([True, True, False, False]),
        gp_params=GPParams(np.float32(-7.9), np.float32(0.1), np.array([0.2, -0.3], np.float32)))
    path = tmp_path / "state.npz"
    bayex.save_state(path, st)
    back = bayex.load_state(path)
    assert list(back.params) == ["a", "k"] and back.params["k"].dtype == np.int32 and back.params["a"].dtype == np.float32
    for k in st.params:
        assert np.array_equal(back.params[k], st.params[k]) and back.best_params[k] == st.best_params[k]
        assert back.best_params[k].dtype == st.best_params[k].dtype
    assert np.array_equal(back.ys, st.ys) and np.array_equal(back.mask, st.mask) and back.mask.dtype == bool
    assert back.best_score == st.best_score
    for a, b in zip(back.gp_params, st.gp_params):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    # the posterior cache key is a digest of exactly these arrays: a reloaded state must hit the same key
    opt = bayex.Optimizer({"a": Real(0.0, 2.0), "k": Integer(-5, 5)}, acq="EI")
    from bayex_b200 import engine
    dig = lambda s: opt._digest(*opt._arrays(s.params, s.ys, s.mask), engine.raw_vector(s.gp_params, 2))
    assert dig(st) == dig(back)


def test_acq_param_passthrough_keeps_reference_defaults():
    dom = {"x": Real(0.0, 1.0)}
    assert bayex.Optimizer(dom, acq="EI").acq_param == np.float32(0.01) or bayex.Optimizer(dom, acq="EI").acq_param == 0.01
    assert bayex.Optimizer(dom, acq="LCB").acq_param == 2.576
    assert bayex.Optimizer(dom, acq="UCB", acq_param=1.5).acq_param == 1.5
    assert bayex.Optimizer(dom, "PI", True).sign == 1  # positional order of the reference is untouched


def test_restart_sweep_concurrency_policy_and_validation():
    """fit_sweep: restarts in flight per GPU by problem size; a bad ``concurrent`` is rejected before any device work."""
    import numpy as np
    import pytest
    from bayex_b200 import _lib, fit_sweep
    assert [fit_sweep.concurrent_restarts(n) for n in (100, 1024, 1025, 2048, 4096, 8192)] == [32, 32, 16, 16, 8, 8]
    assert all(1 <= fit_sweep.concurrent_restarts(n) <= _lib.FIT_BATCH_MAX for n in (1, 10 ** 6))
    x, y, raws = np.zeros((10, 2), np.float32), np.zeros(10, np.float32), np.zeros((3, 4), np.float32)
    for bad in (0, _lib.FIT_BATCH_MAX + 1):
        with pytest.raises(ValueError):
            fit_sweep.fit_restarts(x, y, raws, concurrent=bad)
    with pytest.raises(AssertionError):
        fit_sweep.fit_restarts(x, y, np.zeros((3, 5), np.float32))


def test_weighted_shards_cover_the_draw_exactly():
    """Throughput-weighted candidate shards (multi-GPU load balance): contiguous, ordered, exact cover, on pair-tile
    boundaries; equal shards for small draws or without weights; a faster rank gets more."""
    size, world = (1 << 24) + 12345, 8
    w = [1.0, 1.05, 0.95, 1.0, 1.02, 0.98, 1.0, 1.0]
    r = [bopt.shard_range(size, k, world, w) for k in range(world)]
    assert r[0][0] == 0 and r[-1][0] + r[-1][1] == size
    assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(world - 1))
    assert all(b % bopt.SHARD_GRANULE == 0 for b, _ in r)
    assert r[1][1] > r[0][1] > r[2][1]
    assert [bopt.shard_range(size, k, world) for k in range(world)] == [bopt.shard_range(size, k, world, None) for k in range(world)]
    small = [bopt.shard_range(1000, k, world, w) for k in range(world)]
    assert small == [bopt.shard_range(1000, k, world) for k in range(world)]
    lop = [bopt.shard_range(1 << 20, k, 2, [1e-9, 1.0]) for k in range(2)]  # a rank may end up with (almost) nothing
    assert lop[0][1] + lop[1][1] == 1 << 20 and lop[0][1] <= bopt.SHARD_GRANULE


def test_optimize_suggestion_with_a_generic_callable():
    """optimizer.py:39-73 accepts any scalar function of a (d,) point; a callable that is not a bayex acquisition is ascended
    on the host with central differences under the device refinement's rule: never worse than the start, float32 (d,) out,
    iterates clipped to +-1e6, no GPU involved."""
    from bayex_b200.optimizer import _optimize_suggestion
    c = np.array([0.3, -1.2, 2.0])
    fun = lambda x: -float(np.sum((np.asarray(x, np.float64) - c) ** 2))
    x0 = np.zeros(3, np.float32)
    x1 = _optimize_suggestion(x0, fun, max_iter=10)
    assert x1.shape == (3,) and x1.dtype == np.float32
    assert fun(x1) >= fun(x0) and np.linalg.norm(x1 - c) < 0.05 * np.linalg.norm(x0 - c)
    flat = _optimize_suggestion(x0, lambda x: 1.0, max_iter=10)          # zero gradient: stays put
    assert np.array_equal(flat, x0)
    far = _optimize_suggestion(np.full(2, 9.9e5, np.float32), lambda x: float(np.sum(x)), max_iter=10)
    assert np.all(far <= 1e6) and np.all(far >= 9.9e5)                   # optimizer.py:68
    import pytest
    with pytest.raises(TypeError):
        _optimize_suggestion(x0, lambda x: np.asarray(x), max_iter=2)    # not a scalar function
    with pytest.raises(TypeError):
        _optimize_suggestion(x0, 3.0)
