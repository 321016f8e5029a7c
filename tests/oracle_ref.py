"""TEST INFRASTRUCTURE: drives the reference's own CPU engine (oracle/_ref).

oracle/_ref/libspral_cpu_ref.so is the UNMODIFIED reference C++ CPU numeric
engine (src/ssids/cpu/**) compiled by oracle/Makefile.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; it is the checker, never the product.
"""
import ctypes as C
import os
import sys
import time
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
from spral_b200._lib import Options, Stats, Contrib  # noqa: E402  (struct layouts only)

REF_PATH = os.path.join(_ROOT, "oracle", "_ref", "libspral_cpu_ref.so")
_ref = None


def available():
    return os.path.exists(REF_PATH)


def ensure_env():
    """The reference needs OMP_CANCELLATION before libgomp initialises
    (src/ssids/ssids.f90:1448-1452)."""
    os.environ.setdefault("OMP_CANCELLATION", "TRUE")

def load():
    global _ref
    if _ref is not None:
        return _ref
    ensure_env()
    lib = C.CDLL(REF_PATH, mode=C.RTLD_GLOBAL)
    vp, i, b = C.c_void_p, C.c_int, C.c_bool
    lib.spral_ssids_cpu_create_symbolic_subtree.restype = vp
    lib.spral_ssids_cpu_create_symbolic_subtree.argtypes = [i, i, i, vp, vp, vp, vp, vp, vp, i, vp, C.POINTER(Options)]
    lib.spral_ssids_cpu_destroy_symbolic_subtree.argtypes = [vp]
    lib.oracle_factor.restype = vp
    lib.oracle_factor.argtypes = [b, vp, vp, vp, vp, C.POINTER(Options), vp, i]
    lib.spral_ssids_cpu_destroy_num_subtree_dbl.argtypes = [b, vp]
    for nm in ("fwd", "diag", "diag_bwd", "bwd"):
        f = getattr(lib, f"spral_ssids_cpu_subtree_solve_{nm}_dbl")
        f.restype = i
        f.argtypes = [b, vp, i, vp, i]
    lib.spral_ssids_cpu_subtree_enquire_dbl.argtypes = [b, vp, vp, vp]
    lib.spral_ssids_cpu_subtree_alter_dbl.argtypes = [b, vp, vp]
    lib.spral_ssids_cpu_subtree_get_contrib_dbl.argtypes = [b, vp] + [vp] * 8
    lib.spral_ssids_cpu_subtree_free_contrib_dbl.argtypes = [b, vp]
    lib.oracle_cancellation_enabled.restype = i
    lib.oracle_max_threads.restype = i
    lib.oracle_set_gpu_free_contrib.argtypes = [vp]
    _ref = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefStats(C.Structure):
    """ThreadStats (src/ssids/cpu/ThreadStats.hxx:48-62)."""
    _fields_ = Stats._fields_[:-1]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RefSubtree:
    """One part factorised by the reference CPU engine."""

    def __init__(self, analysis, part, posdef, val, child_contrib=(), options=None,
                 scaling=None, nthreads=0):
        lib = load()
        a = analysis
        self.a, self.posdef = a, bool(posdef)
        self.options = options or Options.default()
        sa, en = int(a.part[part]), int(a.part[part + 1])
        self.cdest = a.part_contrib_dest(part)
        self.symb = lib.spral_ssids_cpu_create_symbolic_subtree(
            a.n, sa, en, _p(a.sptr), _p(a.sparent), _p(a.rptr), _p(a.rlist), _p(a.nptr),
            _p(a.nlist), len(self.cdest), _p(self.cdest), C.byref(self.options))
        self.stats = RefStats()
        self._contribs = list(child_contrib)
        arr = (C.c_void_p * max(1, len(self._contribs)))()
        for k, c in enumerate(self._contribs):
            arr[k] = C.addressof(c)
        self._val = np.ascontiguousarray(val, dtype=np.float64)
        self._sc = None if scaling is None else np.ascontiguousarray(scaling, dtype=np.float64)
        t0 = time.perf_counter()
        self.h = lib.oracle_factor(self.posdef, self.symb, _p(self._val), _p(self._sc),
                                   C.cast(arr, C.c_void_p), C.byref(self.options),
                                   C.byref(self.stats), nthreads)
        self.factor_time = time.perf_counter() - t0

    def solve(self, which, x2, nrhs):
        f = getattr(load(), f"spral_ssids_cpu_subtree_solve_{which}_dbl")
        rc = f(self.posdef, self.h, nrhs, _p(x2), self.a.n)
        assert rc == 0

    def enquire(self):
        n = self.a.n
        if self.posdef:
            d = np.zeros(n)
            load().spral_ssids_cpu_subtree_enquire_dbl(True, self.h, None, _p(d))
            return None, d
        piv = np.zeros(n, dtype=np.int32)
        d = np.zeros(2 * n)
        load().spral_ssids_cpu_subtree_enquire_dbl(False, self.h, _p(piv), _p(d))
        return piv, d

    def get_contrib(self):
        lib = load()
        n, ldval, ndelay, lddelay = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        val, rl, dp, dv = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        lib.spral_ssids_cpu_subtree_get_contrib_dbl(
            self.posdef, self.h, C.cast(C.byref(n), C.c_void_p), C.cast(C.byref(val), C.c_void_p),
            C.cast(C.byref(ldval), C.c_void_p), C.cast(C.byref(rl), C.c_void_p),
            C.cast(C.byref(ndelay), C.c_void_p), C.cast(C.byref(dp), C.c_void_p),
            C.cast(C.byref(dv), C.c_void_p), C.cast(C.byref(lddelay), C.c_void_p))
        c = Contrib(ready=1, n=n.value, val=val.value, ldval=ldval.value, rlist=rl.value,
                    ndelay=ndelay.value, delay_perm=dp.value, delay_val=dv.value,
                    lddelay=lddelay.value, owner=0, posdef=self.posdef, owner_ptr=self.h,
                    device=-1)
        return c

    def close(self):
        lib = load()
        if self.h:
            lib.spral_ssids_cpu_destroy_num_subtree_dbl(self.posdef, self.h)
            self.h = None
        if self.symb:
            lib.spral_ssids_cpu_destroy_symbolic_subtree(self.symb)
            self.symb = None


def ref_factor(analysis, posdef, val, options=None, scaling=None, nthreads=0):
    """Reference fkeep%inner_factor over all parts, sequentially
    (src/ssids/fkeep.F90:61-232).  Returns (parts, inform dict)."""
    a = analysis
    sc = None
    if scaling is not None:
        sc = np.ascontiguousarray(np.asarray(scaling, dtype=np.float64)[a.invp - 1])
    slots = [None] * (a.nparts + 1)
    parts = []
    inform = dict(flag=0, num_delay=0, num_factor=0, num_flops=0, num_neg=0, num_two=0,
                  maxfront=0, maxsupernode=0, matrix_rank=int(a.sptr[a.nnodes]) - 1,
                  not_first_pass=0, not_second_pass=0, factor_time=0.0)
    for p in range(a.nparts):
        lo, hi = int(a.contrib_ptr[p]) - 1, int(a.contrib_ptr[p + 1]) - 1
        cc = [slots[i] for i in range(lo, hi)]
        st = RefSubtree(a, p, posdef, val, cc, options, sc, nthreads)
        parts.append(st)
        s = st.stats
        inform["factor_time"] += st.factor_time
        if s.flag < 0:
            inform["flag"] = min(inform["flag"], s.flag)
            break
        inform["flag"] = max(inform["flag"], s.flag)
        for k in ("num_delay", "num_factor", "num_flops", "num_neg", "num_two",
                  "not_first_pass", "not_second_pass"):
            inform[k] += getattr(s, k)
        inform["maxfront"] = max(inform["maxfront"], s.maxfront)
        inform["maxsupernode"] = max(inform["maxsupernode"], s.maxsupernode)
        inform["matrix_rank"] -= s.num_zero
        idx = int(a.contrib_idx[p]) - 1
        if idx < a.nparts:
            slots[idx] = st.get_contrib()
    return parts, inform, sc


def ref_solve(analysis, parts, posdef, x, scaling_perm=None, job=0):
    """Reference inner_solve_cpu (src/ssids/fkeep.F90:234-323)."""
    a = analysis
    x = np.asarray(x, dtype=np.float64)
    one = x.ndim == 1
    X = np.asfortranarray(x.reshape(a.n, -1))
    nrhs = X.shape[1]
    x2 = np.asfortranarray(X[a.invp - 1, :])
    if scaling_perm is not None and job in (0, 1):
        x2 *= scaling_perm[:, None]
    if job in (0, 1):
        for st in parts:
            st.solve("fwd", x2, nrhs)
    if job == 2:
        for st in parts:
            st.solve("diag", x2, nrhs)
    if job == 3:
        for st in reversed(parts):
            st.solve("bwd", x2, nrhs)
    if job in (0, 4):
        for st in reversed(parts):
            st.solve("diag_bwd", x2, nrhs)
    if scaling_perm is not None and job in (0, 3, 4):
        x2 *= scaling_perm[:, None]
    out = np.empty_like(X)
    out[a.invp - 1, :] = x2
    return out[:, 0] if one else out


def backward_error(A, x, b):
    """Scaled residual of driver/spral_ssids.F90:419-480:
    max_i |r_i| / (||A||_inf ||x||_inf + ||b||_inf), per right-hand side."""
    x = np.asarray(x).reshape(A.shape[0], -1)
    b = np.asarray(b).reshape(A.shape[0], -1)
    r = A @ x - b
    anorm = abs(A).sum(axis=1).max()
    return (np.abs(r).max(axis=0) / (anorm * np.abs(x).max(axis=0) + np.abs(b).max(axis=0))).max()
