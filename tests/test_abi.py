"""The C-ABI library loads (no GPU needed) and exports every symbol include/bayex_b200.h declares."""
import ctypes
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "bayex_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bayex_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_error_string():
    from bayex_b200 import _lib
    assert _lib.lib.bx_version() == 100
    assert isinstance(_lib.lib.bx_last_error(), bytes)


def test_packed_argmax_key_semantics():
    """max over packed keys == jnp.argmax: largest value, NaN beats everything, first index among ties."""
    from bayex_b200 import _lib
    pack = _lib.lib.bx_pack
    vals = [-np.inf, -3.5, -0.0, 0.0, 1e-30, 2.0, np.inf]
    keys = [pack(ctypes.c_float(v), 7) for v in vals]
    assert keys[2] == keys[3], "-0.0 and +0.0 compare equal"
    assert keys == sorted(keys)
    assert pack(ctypes.c_float(float("nan")), 9) > keys[-1]
    assert pack(ctypes.c_float(1.0), 3) > pack(ctypes.c_float(1.0), 4)  # lower index wins ties
    for v, i in [(1.5, 0), (-2.25, 123456), (0.0, 2 ** 32 - 1)]:
        got_v, got_i = _lib.unpack(pack(ctypes.c_float(v), i))
        assert got_v == v and got_i == i
    v, _ = _lib.unpack(pack(ctypes.c_float(float("nan")), 1))
    assert math.isnan(v)


def test_bad_arguments_are_rejected_without_a_gpu():
    from bayex_b200 import _lib
    assert _lib.lib.bx_fit_workspace_bytes(0, 3) == 0
    assert _lib.lib.bx_fit_workspace_bytes(100, 3) > 2 * 128 * 128 * 4
    rc = _lib.lib.bx_topk(None, 10, 5, 0, None, None, None, 0, None)
    assert rc == -1 and b"null" in _lib.lib.bx_last_error()


def test_batched_fit_rejects_bad_arguments_without_a_gpu():
    """bx_fit_adamw_batch validates before touching the device: null pointers, R outside [1, BX_FIT_BATCH_MAX], a
    workspace smaller than R slots."""
    import ctypes
    from bayex_b200 import _lib
    f = _lib.lib.bx_fit_adamw_batch
    fake = ctypes.c_void_p(256)  # never dereferenced: every call below fails validation first
    assert f(None, fake, 100, 3, fake, 4, 1e-3, 10, None, fake, 1 << 30, None) == -1
    assert f(fake, fake, 100, 3, fake, 0, 1e-3, 10, None, fake, 1 << 30, None) == -2
    assert f(fake, fake, 100, 3, fake, _lib.FIT_BATCH_MAX + 1, 1e-3, 10, None, fake, 1 << 30, None) == -2
    slot = _lib.lib.bx_fit_workspace_bytes(100, 3)
    assert f(fake, fake, 100, 3, fake, 4, 1e-3, 10, None, fake, 4 * slot - 1, None) == -3
    assert b"workspace" in _lib.lib.bx_last_error()


def test_factor_append_rejects_bad_arguments_without_a_gpu():
    """bx_factor_append validates before touching the device: null pointers, a factor without a free padding row (the caller
    must rebuild), a workspace that is too small."""
    import ctypes
    from bayex_b200 import _lib
    f, ws = _lib.lib.bx_factor_append, _lib.lib.bx_factor_append_workspace_bytes
    assert ws(0) == 0 and ws(100) == 0 and ws(128) >= 8 * 128 * 4
    fake = 256  # never dereferenced: every call below fails validation first
    fac = lambda n, np_: _lib.FactorT(n, np_, 3, 4, fake, fake, 0, 0, fake, fake, fake)
    assert f(ctypes.byref(fac(100, 128)), None, fake, fake, 1 << 20, None) == _lib.E_ARG
    assert f(ctypes.byref(fac(128, 128)), fake, fake, fake, 1 << 20, None) == _lib.E_RANGE
    assert b"no free row" in _lib.lib.bx_last_error()
    assert f(ctypes.byref(fac(100, 128)), fake, fake, fake, ws(128) - 1, None) == _lib.E_WORKSPACE
    assert f(ctypes.byref(_lib.FactorT(100, 128, 3, 8, fake, fake, 0, 0, fake, fake, fake)), fake, fake, fake, 1 << 20, None) == _lib.E_ARG
