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


def _p(a):
    """numpy array -> void* (a bare Python int would be passed as a 32-bit C int: no argtypes on this handle)."""
    return None if a is None else C.c_void_p(a.ctypes.data)


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
    lib.spral_ssids_analyse_coord(5, _p(order), C.c_int64(len(ROW)),
                                  _p(ROW), _p(COL), _p(val),
                                  C.byref(akeep), C.byref(opt), C.byref(inf))
    return akeep, opt, inf


def test_coordinate_input_is_cleaned_like_the_reference(capi):
    akeep, opt, inf = _coord(capi, 1, None)
    assert inf.flag == 3 and inf.matrix_dup == 1 and inf.matrix_outrange == 1 and inf.matrix_missing_diag == 0
    assert inf.num_factor == 15 and inf.num_flops == 55 and inf.matrix_rank == 5      # examples/C/ssids.c
    fkeep = C.c_void_p(None)
    inf2 = Inform()
    capi.spral_ssids_factor(False, None, None, _p(VAL), None, akeep, C.byref(fkeep), C.byref(opt), C.byref(inf2))
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


def _csc(lib, ptr, row, n, check=True, base=1, order=None, ordering=1):
    opt, inf = Options(), Inform()
    lib.spral_ssids_default_options(C.byref(opt))
    opt.array_base, opt.ordering = base, ordering
    ptr = np.asarray(ptr, dtype=np.int64)
    row = np.asarray(row, dtype=np.int32)
    akeep = C.c_void_p(None)
    lib.spral_ssids_analyse(check, n, _p(order), _p(ptr),
                            _p(row), None, C.byref(akeep), C.byref(opt), C.byref(inf))
    lib.spral_ssids_free_akeep(C.byref(akeep))
    return inf


def test_csc_data_checking_warnings_and_errors(capi):
    """The warnings / errors of ssids_analyse(check = true) (src/ssids/datatypes.f90:25-59; the reference's
    tests/ssids/ssids.f90 test_warnings / test_errors): duplicates (2), out-of-range (1), both (3), missing
    diagonal (4), missing diagonal with duplicates or out-of-range (5), all out of range (-4), bad ptr (-3),
    n < 0 (-2)."""
    # 3 x 3 tridiagonal, lower triangle, 1-based: the clean case
    inf = _csc(capi, [1, 3, 5, 6], [1, 2, 2, 3, 3], 3)
    assert inf.flag == 0 and inf.matrix_dup == 0 and inf.matrix_outrange == 0 and inf.matrix_missing_diag == 0
    assert inf.matrix_rank == 3 and inf.num_factor == 6      # nemin = 32 merges the three columns: dense 3 x 3
    inf = _csc(capi, [1, 4, 6, 7], [1, 2, 2, 2, 3, 3], 3)                  # (2,1) twice
    assert inf.flag == 2 and inf.matrix_dup == 1
    inf = _csc(capi, [1, 4, 6, 7], [1, 2, 9, 2, 3, 3], 3)                  # row 9 out of range
    assert inf.flag == 1 and inf.matrix_outrange == 1
    inf = _csc(capi, [1, 5, 7, 8], [1, 2, 2, 9, 2, 3, 3], 3)
    assert inf.flag == 3 and inf.matrix_dup == 1 and inf.matrix_outrange == 1
    inf = _csc(capi, [1, 3, 4, 5], [1, 2, 3, 3], 3)                        # (2,2) absent
    assert inf.flag == 4 and inf.matrix_missing_diag == 1
    inf = _csc(capi, [1, 4, 5, 6], [1, 2, 2, 3, 3], 3)                     # (2,2) absent and a duplicate
    assert inf.flag == 5
    inf = _csc(capi, [1, 3, 5, 6], [7, 8, 9, 7, 8], 3)
    assert inf.flag == -4
    inf = _csc(capi, [1, 3, 2, 6], [1, 2, 2, 3, 3], 3)
    assert inf.flag == -3
    inf = _csc(capi, [1], [], -1)
    assert inf.flag == -2
    # entries given in the UPPER triangle, 0-based: mirrored, same prediction as the clean case
    inf = _csc(capi, [0, 1, 3, 5], [0, 0, 1, 1, 2], 3, base=0)
    assert inf.flag == 0 and inf.num_factor == 6
    # structurally singular (an empty row and column): warning 6 unless data warnings take over
    inf = _csc(capi, [1, 2, 2, 3], [1, 3], 3, check=False)
    assert inf.flag == 6 and inf.matrix_rank == 2


# simple_mat_lower of the reference's tests (tests/ssids/ssids.f90:1475-1510)
SM_PTR = [1, 4, 5, 7, 8]
SM_ROW = [1, 2, 4, 2, 3, 4, 4]
SM_COL = [1, 1, 1, 2, 3, 3, 4]


def _analyse_raw(lib, n, ptr, row, ordering, order=None, val=None, check=True, nemin=8):
    opt, inf = Options(), Inform()
    lib.spral_ssids_default_options(C.byref(opt))
    opt.array_base, opt.ordering, opt.nemin = 1, ordering, nemin
    ptr = np.asarray(ptr, dtype=np.int64)
    row = np.asarray(row, dtype=np.int32)
    akeep = C.c_void_p(None)
    lib.spral_ssids_analyse(check, n, _p(order), _p(ptr),
                            _p(row), _p(val), C.byref(akeep),
                            C.byref(opt), C.byref(inf))
    lib.spral_ssids_free_akeep(C.byref(akeep))
    return inf.flag


def test_the_reference_test_errors_of_analyse(capi):
    """test_errors of tests/ssids/ssids.f90:140-345, the calls that end inside ssids_analyse /
    ssids_analyse_coord, with the flags the reference expects."""
    ident = np.arange(1, 5, dtype=np.int32)
    assert _analyse_raw(capi, -1, SM_PTR, SM_ROW, 0, ident) == -2                       # n < 0
    assert _analyse_raw(capi, 4, [0, 4, 5, 7, 8], SM_ROW, 0, ident) == -3               # ptr with zero component
    assert _analyse_raw(capi, 4, [1, 5, 4, 7, 8], SM_ROW, 0, ident) == -3               # non-monotonic ptr
    assert _analyse_raw(capi, 4, SM_PTR, [0, 0, 0, 0, 0, 0, 0], 0, ident) == -4         # all of A%row out of range
    assert _analyse_raw(capi, 4, SM_PTR, SM_ROW, 0, ident, nemin=-1) == 0               # nemin oor -> default, success
    assert _analyse_raw(capi, 4, SM_PTR, SM_ROW, 0, None) == -8                         # order absent
    assert _analyse_raw(capi, 4, SM_PTR, SM_ROW, 0, np.array([5, 2, 3, 4], np.int32)) == -8    # out of range above
    assert _analyse_raw(capi, 4, SM_PTR, SM_ROW, 0, np.array([0, 2, 3, 4], np.int32)) == -8    # out of range below
    assert _analyse_raw(capi, 4, SM_PTR, SM_ROW, 0, np.array([1, 1, 1, 1], np.int32)) == -8    # not a permutation
    assert _analyse_raw(capi, 4, SM_PTR, SM_ROW, -1, ident) == -8                       # options%ordering out of range
    assert _analyse_raw(capi, 4, SM_PTR, SM_ROW, 3, ident) == -8
    assert _analyse_raw(capi, 4, SM_PTR, SM_ROW, 2, ident) == -9                        # val absent
    # coordinate form
    row, col = np.array(SM_ROW, np.int32), np.array(SM_COL, np.int32)

    def coord(n, ne, ordering, order=None):
        opt, inf = Options(), Inform()
        capi.spral_ssids_default_options(C.byref(opt))
        opt.array_base, opt.ordering = 1, ordering
        akeep = C.c_void_p(None)
        capi.spral_ssids_analyse_coord(n, _p(order), C.c_int64(ne), _p(row),
                                       _p(col), None, C.byref(akeep), C.byref(opt), C.byref(inf))
        capi.spral_ssids_free_akeep(C.byref(akeep))
        return inf.flag
    assert coord(4, 7, 0, np.array([5, 2, 3, 4], np.int32)) == -8
    assert coord(4, 7, 0, np.array([0, 2, 3, 4], np.int32)) == -8
    assert coord(4, 7, 25, ident.copy()) == -8
    assert coord(4, 7, 2, ident.copy()) == -9
    assert coord(4, 7, 0, None) == -8
    assert coord(-1, 7, 0, ident.copy()) == -2                                          # n < 0
    assert coord(4, -1, 0, ident.copy()) == -4                                          # ne < 0
    assert coord(4, 7, 1) == 0
