"""Scaling pre-processing of ssids_factor (options%scaling), host side.

Mirrors the reference's spral_scaling entry points that SSIDS uses
(src/ssids/ssids.f90:861-1028 dispatch; src/scaling.f90): `hungarian_scale_sym` (MC64-type
matching-based scaling, options%scaling = 1) and `equilib_scale_sym` (infinity-norm
equilibration, options%scaling = 4).  Both return the vector that `factor(...,
scaling=...)` applies as S A S.  The arithmetic is in csrc/scaling.cpp."""
import ctypes as C

import numpy as np

from . import _lib

WARNING_SINGULAR = 1
ERROR_SINGULAR = -2


def _args(n, ptr, row, val):
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    row = np.ascontiguousarray(row, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    assert len(ptr) == n + 1 and len(row) >= ptr[n] - 1 and len(val) >= ptr[n] - 1
    return ptr, row, val


def hungarian_scale_sym(n, ptr, row, val, scale_if_singular=False):
    """(scaling, match, flag, matched): src/scaling.f90:134-170.  match[i] is the 1-based column
    matched to row i+1 (negative outside the matching of a structurally singular matrix)."""
    ptr, row, val = _args(n, ptr, row, val)
    scaling = np.empty(n)
    match = np.empty(n, dtype=np.int32)
    matched = C.c_int(0)
    flag = _lib.load().spral_ssids_b200_hungarian_scale_sym(
        n, ptr.ctypes.data, row.ctypes.data, val.ctypes.data, scaling.ctypes.data, match.ctypes.data,
        1 if scale_if_singular else 0, C.byref(matched))
    return scaling, match, flag, matched.value


def equilib_scale_sym(n, ptr, row, val, max_iterations=10, tol=1e-8):
    """(scaling, iterations): src/scaling.f90:480-521 (Knight, Ruiz, Ucar, Algorithm 1)."""
    ptr, row, val = _args(n, ptr, row, val)
    scaling = np.empty(n)
    it = C.c_int(0)
    _lib.load().spral_ssids_b200_equilib_scale_sym(n, ptr.ctypes.data, row.ctypes.data, val.ctypes.data,
                                                   scaling.ctypes.data, max_iterations, tol, C.byref(it))
    return scaling, it.value


def match_order_metis(n, ptr, row, val):
    """(order, scaling, flag): src/match_order.f90:51-208, options%ordering = 2.  order[i] is the 1-based
    pivot position of variable i+1 (pass it to analyse(order=...)); matched pairs are consecutive."""
    ptr, row, val = _args(n, ptr, row, val)
    order = np.zeros(n, dtype=np.int32)
    scaling = np.empty(n)
    flag = _lib.load().spral_ssids_b200_match_order_metis(n, ptr.ctypes.data, row.ctypes.data, val.ctypes.data,
                                                          order.ctypes.data, scaling.ctypes.data)
    if flag < 0:
        raise RuntimeError(f"match_order_metis failed with flag {flag}")
    return order, scaling, flag


def auction_scale_sym(n, ptr, row, val):
    """(scaling, match, matched, iterations): src/scaling.f90:269-309 with the default auction_options;
    options%scaling = 2.  An approximate matching: a few rows may stay unmatched (match[i] = 0)."""
    ptr, row, val = _args(n, ptr, row, val)
    scaling = np.empty(n)
    match = np.zeros(n, dtype=np.int32)
    matched, it = C.c_int(0), C.c_int(0)
    _lib.load().spral_ssids_b200_auction_scale_sym(n, ptr.ctypes.data, row.ctypes.data, val.ctypes.data,
                                                   scaling.ctypes.data, match.ctypes.data, None,
                                                   C.byref(matched), C.byref(it))
    return scaling, match, matched.value, it.value
