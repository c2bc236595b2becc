"""TEST INFRASTRUCTURE: Python handle on the CPU oracle (oracle/liboracle.so, prefix orc_).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under minimaloptix_b200/ does.
"""
import ctypes as C
import os

from minimaloptix_b200 import structs as S
from minimaloptix_b200._binding import Backend

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "liboracle.so")

_vp, _u32, _i32p, _f3 = C.c_void_p, C.c_uint32, C.POINTER(C.c_int32), C.c_float * 3
_EXTRA = {
    "set_threads": (C.c_int, [_vp, C.c_int]),
    "set_brute_force": (C.c_int, [_vp, C.c_int]),
    "set_quad_light_draw_order": (C.c_int, [_vp, C.c_int]),
    "tea16": (_u32, [_u32, _u32]),
    "lcg": (_u32, [_i32p]),
    "rand": (C.c_float, [_i32p]),
    "philox": (None, [_u32 * 4, _u32 * 2, _u32 * 4]),
    "disney_eval": (None, [C.POINTER(S.DisneyParams), _f3, _f3, _f3, _f3, _f3, _f3]),
    "disney_pdf": (C.c_float, [C.POINTER(S.DisneyParams), _f3, _f3, _f3, _f3]),
    "disney_sample": (None, [_i32p, C.POINTER(S.DisneyParams), _f3, _f3, _f3, _f3]),
    "refine_hitpoint": (None, [_f3, _f3, _f3, _f3, _f3, _f3]),
    "refract": (C.c_int, [_f3, _f3, C.c_float, _f3]),
    "fresnel": (C.c_float, [C.c_float, C.c_float, C.c_float]),
}

_backend = None


def backend():
    global _backend
    if _backend is None:
        _backend = Backend(ORACLE_LIB, "orc_", extra=_EXTRA)
    return _backend


def context(threads=0, brute_force=False):
    b = backend()
    ctx = b.context(0)
    b.set_threads(ctx.h, threads)
    b.set_brute_force(ctx.h, 1 if brute_force else 0)
    return ctx
