"""ctypes binding of ``libbayex_b200.so`` (the C ABI declared in ``include/bayex_b200.h``).

There is no CPU fallback: importing this module raises if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``bayex_b200/csrc/build.sh``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BAYEX_B200_LIB") or os.path.join(HERE, "libbayex_b200.so")  # override: tuning variants

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: bayex_b200 has no CPU fallback. Build the CUDA library first "
        "(bayex_b200/csrc/build.sh, or __graft_entry__.build()).")

lib = C.CDLL(LIB_PATH)

ACQ_IDS = {"EI": 0, "PI": 1, "UCB": 2, "LCB": 3}
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05_F16, ENGINE_TCGEN05_F16_RAW, ENGINE_TCGEN05_F16_ACC = 0, 1, 2, 3, 4
MAX_D, MAX_K, TILE, FIT_BATCH_MAX = 64, 1024, 128, 32
E_ARG, E_RANGE, E_WORKSPACE, E_NOTPD, E_UNSUPPORTED = -1, -2, -3, -4, -5
POINTS_TC_MIN = 65536  # BX_POINTS_TC_MIN


class DimT(C.Structure):
    _fields_ = [("kind", C.c_int32), ("lo", C.c_float), ("hi", C.c_float), ("ilo", C.c_int32), ("ihi", C.c_int32)]


class FactorT(C.Structure):
    _fields_ = [("n", C.c_int32), ("np", C.c_int32), ("d", C.c_int32), ("d4", C.c_int32),
                ("d_U", C.c_void_p), ("d_T", C.c_void_p),
                ("d_T16hi", C.c_void_p), ("d_T16lo", C.c_void_p), ("d_alpha", C.c_void_p), ("d_scale", C.c_void_p), ("d_hyp", C.c_void_p)]


_vp, _i32, _u64, _f32, _sz = C.c_void_p, C.c_int32, C.c_uint64, C.c_float, C.c_size_t
_key = C.POINTER(C.c_uint32)

# name -> (restype, argtypes); must list every symbol include/bayex_b200.h declares (tests/test_abi.py checks).
SIGNATURES = {
    "bx_version": (_i32, []),
    "bx_last_error": (C.c_char_p, []),
    "bx_check_status": (_i32, [_vp, _vp]),
    "bx_threefry_fill": (_i32, [_key, C.POINTER(DimT), _i32, _u64, _u64, _vp, _vp]),
    "bx_domain_sample": (_i32, [_key, C.POINTER(DimT), _u64, _u64, _vp, _vp]),
    "bx_threefry_gather": (_i32, [_key, C.POINTER(DimT), _i32, _vp, _i32, _vp, _vp]),
    "bx_cov": (_i32, [_vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "bx_fit_workspace_bytes": (_sz, [_i32, _i32]),
    "bx_nlml_grad": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "bx_fit_adamw": (_i32, [_vp, _vp, _i32, _i32, _vp, _f32, _i32, _vp, _vp, _sz, _i32, _vp]),
    "bx_fit_adamw_batch": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _f32, _i32, _vp, _vp, _sz, _vp]),
    "bx_factor_build": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bx_factor_append_workspace_bytes": (_sz, [_i32]),
    "bx_factor_append": (_i32, [C.POINTER(FactorT), _vp, _vp, _vp, _sz, _vp]),
    "bx_score_workspace_bytes": (_sz, [_u64]),
    "bx_score_points": (_i32, [C.POINTER(FactorT), _vp, _u64, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _u64, _vp, _sz, _vp]),
    "bx_score_generated": (_i32, [C.POINTER(FactorT), _key, C.POINTER(DimT), _u64, _u64, _i32, _f32, _f32, _vp, _vp,
                                  _i32, _vp, _sz, _vp]),
    "bx_score_topk_workspace_bytes": (_sz, [_u64, _i32]),
    "bx_score_generated_topk": (_i32, [C.POINTER(FactorT), _key, C.POINTER(DimT), _u64, _u64, _i32, _f32, _f32, _i32, _vp,
                                       _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "bx_topk_merge": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bx_propose": (_i32, [_key, C.POINTER(DimT), _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
    "bx_topk_workspace_bytes": (_sz, [_u64, _i32]),
    "bx_topk": (_i32, [_vp, _u64, _i32, _u64, _vp, _vp, _vp, _sz, _vp]),
    "bx_refine_workspace_bytes": (_sz, [C.POINTER(FactorT), _i32]),
    "bx_refine": (_i32, [C.POINTER(FactorT), _vp, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _sz, _vp]),
    "bx_pack": (_u64, [_f32, C.c_uint32]),
    "bx_unpack": (None, [_u64, C.POINTER(_f32), C.POINTER(C.c_uint32)]),
    "bx_selftest_umma": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "bx_selftest_ffma": (_i32, [_vp, _i32, _i32, _vp]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class BayexB200Error(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.bx_last_error().decode("utf-8", "replace")
        kind = "bad argument" if rc < 0 else "CUDA error"
        raise BayexB200Error(f"{what} failed ({kind} {rc}): {msg}")


def key_words(key):
    arr = (C.c_uint32 * 2)(int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF)
    return arr


def unpack(packed: int):
    v, i = _f32(), C.c_uint32()
    lib.bx_unpack(C.c_uint64(packed & 0xFFFFFFFFFFFFFFFF), C.byref(v), C.byref(i))
    return v.value, i.value
