"""CPU checks of the spral_ssids.h-compatible C interface up to the end of analyse (no device: the
SPRAL_B200_ANALYSE_ONLY hook skips the symbolic subtrees): coordinate input and data cleaning
(src/ssids/ssids.f90:392-700, clean_coord / clean_cscl_oop of matrix_util.f90), the orderings of
options%ordering = 0 / 1 / 2, flags and the analyse-time inform of the reference's C example."""
import ctypes as C
import os

import numpy as np
import pytest

from spral_b200 import _lib


class Options(C.Structure):                      # include/spral_ssids_compat.h (== include/spral_ssids.h:15-36)
    _fields_ = [("array_base", C.c_int), ("print_level", C.c_int), ("unit_diagnostics", C.c_int),
                ("unit_error", C.c_int), ("unit_warning", C.c_int), ("ordering", C.c_int), ("nemin", C.c_int),
                ("ignore_numa", C.c_bool), ("use_gpu", C.c_bool), ("min_gpu_work", C.c_int64),
                ("max_load_inbalance", C.c_float), ("gpu_perf_coeff", C.c_float), ("scaling", C.c_int),
                ("small_subtree_threshold", C.c_int64), ("cpu_block_size", C.c_int), ("action", C.c_bool),
                ("pivot_method", C.c_int), ("small", C.c_double), ("u", C.c_double), ("unused", C.c_char * 80)]


class Inform(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("flag", "matrix_dup", "matrix_missing_diag", "matrix_outrange", "matrix_rank",
                                       "maxdepth", "maxfront", "num_delay")] + \
               [("num_factor", C.c_int64), ("num_flops", C.c_int64)] + \
               [(k, C.c_int) for k in ("num_neg", "num_sup", "num_two", "stat", "cuda_error", "cublas_error",
                                       "maxsupernode")] + [("unused", C.c_char * 76)]


@pytest.fixture()
def capi(monkeypatch):
    monkeypatch.setenv("SPRAL_B200_ANALYSE_ONLY", "1")
    lib = C.CDLL(_lib.LIB_PATH)
    assert C.sizeof(Options) == 176 and C.sizeof(Inform) == 152     # the reference's sizes (tests/test_abi.py)
    return lib


ROW = np.array([5, 1, 2, 4, 2, 3, 2, 5, 4, 2, 7], dtype=np.int32)       # 5x5 example, shuffled, 1-based;
COL = np.array([5, 1, 3, 4, 2, 3, 1, 2, 3, 2, 1], dtype=np.int32)       # (2,3) upper, (2,2) twice, (7,1) out of range
VAL = np.array([2.0, 2.0, 1.0, -1.0, 1.5, 3.0, 1.0, 1.0, 2.0, 2.5, 9.0])


def _coord(lib, ordering, val, order=None):
    opt, inf = Options(), Inform()
    lib.spral_ssids_default_options(C.byref(opt))
    opt.array_base, opt.ordering = 1, ordering
    akeep = C.c_void_p(None)
    lib.spral_ssids_analyse_coord(5, order.ctypes.data if order is not None else None, C.c_int64(len(ROW)),
                                  ROW.ctypes.data, COL.ctypes.data, val.ctypes.data if val is not None else None,
                                  C.byref(akeep), C.byref(opt), C.byref(inf))
    return akeep, opt, inf


def test_coordinate_input_is_cleaned_like_the_reference(capi):
    akeep, opt, inf = _coord(capi, 1, None)
    assert inf.flag == 3 and inf.matrix_dup == 1 and inf.matrix_outrange == 1 and inf.matrix_missing_diag == 0
    assert inf.num_factor == 15 and inf.num_flops == 55 and inf.matrix_rank == 5      # examples/C/ssids.c
    fkeep = C.c_void_p(None)
    inf2 = Inform()
    capi.spral_ssids_factor(False, None, None, VAL.ctypes.data, None, akeep, C.byref(fkeep), C.byref(opt), C.byref(inf2))
    assert inf2.flag == -1                                          # analyse-only akeep: call sequence error
    assert capi.spral_ssids_free_akeep(C.byref(akeep)) == 0


def test_orderings(capi):
    order = np.zeros(5, dtype=np.int32)
    akeep, opt, inf = _coord(capi, 2, None, order)                  # matching-based ordering without values
    assert inf.flag == -9
    capi.spral_ssids_free_akeep(C.byref(akeep))
    akeep, opt, inf = _coord(capi, 2, VAL, order)
    assert inf.flag == 3 and sorted(order.tolist()) == [1, 2, 3, 4, 5]
    capi.spral_ssids_free_akeep(C.byref(akeep))
    user = np.array([3, 1, 2, 5, 4], dtype=np.int32)                # user ordering, returned (possibly refined)
    akeep, opt, inf = _coord(capi, 0, None, user)
    assert inf.flag == 3 and sorted(user.tolist()) == [1, 2, 3, 4, 5]
    capi.spral_ssids_free_akeep(C.byref(akeep))
    bad = np.array([1, 1, 2, 3, 4], dtype=np.int32)
    akeep, opt, inf = _coord(capi, 0, None, bad)
    assert inf.flag == -8
    akeep, opt, inf = _coord(capi, 7, None)
    assert inf.flag == -8                                           # options%ordering out of range
