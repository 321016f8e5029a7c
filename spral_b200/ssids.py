"""Host-side mirror of the reference's SSIDS interfaces on top of the C ABI.

Names follow the reference: `analyse` / `factor` / `solve` mirror the
`spral_ssids` module (src/ssids/ssids.f90:33-40); `SymbolicSubtree` /
`NumericSubtree` mirror `symbolic_subtree_base` / `numeric_subtree_base`
(src/ssids/subtree.f90:26-126).  All numerics happen in
libspral_ssids_b200.so on the GPU; this file only marshals arrays.
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import Options, Stats, Contrib, AnalysisView


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Analysis:
    """Result of the symbolic phase (mirrors ssids_akeep, src/ssids/akeep.f90:25-80).

    ngpu: number of GPUs to partition for (topology = one region with ngpu
    GPUs and no CPU resource; the reference's `gpu_only` intent)."""

    def __init__(self, n, ptr, row, order=None, nemin=32, ngpu=1,
                 min_gpu_work=0, max_load_inbalance=1.2, gpu_perf_coeff=1.0):
        lib = _lib.load()
        self.n = n
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        self.row = np.ascontiguousarray(row, dtype=np.int32)
        if order is None:
            order = np.zeros(n, dtype=np.int32)
            rc = lib.spral_ssids_b200_metis_order(n, _ptr(self.ptr), _ptr(self.row), _ptr(order)) if n > 0 else 0
            if rc != 0:
                raise RuntimeError(f"metis_order failed: {rc}")
        self.order = np.ascontiguousarray(order, dtype=np.int32).copy()
        flag = C.c_int(0)
        self._h = lib.spral_ssids_b200_analyse(
            n, _ptr(self.ptr), _ptr(self.row), _ptr(self.order), nemin, -abs(ngpu),
            min_gpu_work, max_load_inbalance, gpu_perf_coeff, C.byref(flag))
        self.flag = flag.value
        self._load_view()

    def set_partition(self, part, exec_loc):
        """Replaces the subtree partition (spral_ssids_b200_analysis_set_partition): part = 1-based first nodes,
        len nparts + 1; exec_loc = the reference's location codes (rank r of an N-GPU job: r + 2)."""
        part = np.ascontiguousarray(part, dtype=np.int32)
        loc = np.ascontiguousarray(exec_loc, dtype=np.int32)
        rc = _lib.load().spral_ssids_b200_analysis_set_partition(self._h, len(loc), _ptr(part), _ptr(loc))
        if rc != 0:
            raise ValueError("invalid subtree partition")
        self._load_view()

    def _load_view(self):
        lib = _lib.load()
        n = self.n
        v = AnalysisView()
        lib.spral_ssids_b200_analysis_get(self._h, C.byref(v))
        self.view = v
        self.nnodes, self.nparts = v.nnodes, v.nparts
        as_np = np.ctypeslib.as_array
        nn = v.nnodes
        self.sptr = as_np(v.sptr, (nn + 1,))
        self.sparent = as_np(v.sparent, (nn,)) if nn else np.zeros(0, np.int32)
        self.rptr = as_np(v.rptr, (nn + 1,))
        self.rlist = as_np(v.rlist, (int(self.rptr[nn]) - 1,)) if nn else np.zeros(0, np.int32)
        self.nptr = as_np(v.nptr, (nn + 1,))
        nz = int(self.ptr[n]) - 1
        self.nlist = as_np(v.nlist, (2 * nz,)) if nz else np.zeros(0, np.int64)
        npt = v.nparts
        self.invp = as_np(v.invp, (n,)) if n else np.zeros(0, np.int32)
        self.part = as_np(v.part, (npt + 1,))
        self.exec_loc = as_np(v.exec_loc, (npt,)) if npt else np.zeros(0, np.int32)
        self.contrib_ptr = as_np(v.contrib_ptr, (npt + 3,))
        self.contrib_idx = as_np(v.contrib_idx, (npt,)) if npt else np.zeros(0, np.int32)
        self.contrib_dest = as_np(v.contrib_dest, (npt,)) if npt else np.zeros(0, np.int32)
        self.num_factor, self.num_flops = v.num_factor, v.num_flops
        self.maxfront, self.maxsupernode, self.maxdepth = v.maxfront, v.maxsupernode, v.maxdepth

    def part_contrib_dest(self, p):
        """contrib_dest(contrib_ptr(p):contrib_ptr(p+1)-1) of part p (0-based p)."""
        a, b = int(self.contrib_ptr[p]) - 1, int(self.contrib_ptr[p + 1]) - 1
        return np.ascontiguousarray(self.contrib_dest[a:b], dtype=np.int32)

    def close(self):
        if self._h:
            _lib.load().spral_ssids_b200_analysis_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SymbolicSubtree:
    """symbolic_subtree_base for one part on one B200
    (construct_gpu_symbolic_subtree, src/ssids/gpu/subtree.f90:76-160)."""

    def __init__(self, analysis, part, device=0, options=None):
        lib = _lib.load()
        self.analysis, self.part, self.device = analysis, part, device
        self.options = options or Options.default()
        a = analysis
        self.sa, self.en = int(a.part[part]), int(a.part[part + 1])
        self.contrib_dest = a.part_contrib_dest(part)
        self._h = lib.spral_ssids_gpu_create_symbolic_subtree(
            device, a.n, self.sa, self.en, _ptr(a.sptr), _ptr(a.sparent), _ptr(a.rptr),
            _ptr(a.rlist), _ptr(a.nptr), _ptr(a.nlist), len(self.contrib_dest),
            _ptr(self.contrib_dest), C.byref(self.options))
        if not self._h:
            raise RuntimeError("spral_ssids_gpu_create_symbolic_subtree failed")

    def maps(self):
        """(rlist_direct, level_ptr, level_list) as built on the device."""
        lib = _lib.load()
        a = self.analysis
        nloc = self.en - self.sa
        nr = int(a.rptr[self.en - 1] - a.rptr[self.sa - 1])
        rd = np.zeros(nr, dtype=np.int32)
        nl = C.c_int(0)
        lptr = np.zeros(nloc + 2, dtype=np.int32)
        llist = np.zeros(nloc, dtype=np.int32)
        lib.spral_ssids_gpu_symbolic_get_maps(self._h, _ptr(rd), C.byref(nl), _ptr(lptr), _ptr(llist))
        return rd, lptr[:nl.value + 1].copy(), llist

    def factor(self, posdef, aval, child_contrib=(), options=None, scaling=None):
        return NumericSubtree(self, posdef, aval, child_contrib, options or self.options, scaling)

    def close(self):
        if self._h:
            _lib.load().spral_ssids_gpu_destroy_symbolic_subtree(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NumericSubtree:
    """numeric_subtree_base: factors of one part, resident in HBM."""

    def __init__(self, symb, posdef, aval, child_contrib, options, scaling):
        lib = _lib.load()
        self.symb, self.posdef = symb, bool(posdef)
        self.stats = Stats()
        self._contribs = list(child_contrib)          # keep alive
        arr = (C.c_void_p * max(1, len(self._contribs)))()
        for i, c in enumerate(self._contribs):
            arr[i] = C.addressof(c)
        aval_p = aval if isinstance(aval, int) else _ptr(np.ascontiguousarray(aval, dtype=np.float64))
        self._aval_keep = aval
        sc_p = None
        if scaling is not None:
            sc_p = scaling if isinstance(scaling, int) else _ptr(np.ascontiguousarray(scaling, dtype=np.float64))
        self._h = lib.spral_ssids_gpu_create_num_subtree_dbl(
            self.posdef, symb._h, aval_p, sc_p, C.cast(arr, C.c_void_p),
            C.byref(options), C.byref(self.stats))

    def _solve(self, fn, x, nrhs, ldx):
        xp = x if isinstance(x, int) else _ptr(x)
        rc = fn(self.posdef, self._h, nrhs, xp, ldx)
        if rc < 0:
            raise RuntimeError(f"solve failed with flag {rc}")

    def solve_fwd(self, x, nrhs=1, ldx=None):
        self._solve(_lib.load().spral_ssids_gpu_subtree_solve_fwd_dbl, x, nrhs, ldx or self.symb.analysis.n)

    def solve_diag(self, x, nrhs=1, ldx=None):
        self._solve(_lib.load().spral_ssids_gpu_subtree_solve_diag_dbl, x, nrhs, ldx or self.symb.analysis.n)

    def solve_diag_bwd(self, x, nrhs=1, ldx=None):
        self._solve(_lib.load().spral_ssids_gpu_subtree_solve_diag_bwd_dbl, x, nrhs, ldx or self.symb.analysis.n)

    def solve_bwd(self, x, nrhs=1, ldx=None):
        self._solve(_lib.load().spral_ssids_gpu_subtree_solve_bwd_dbl, x, nrhs, ldx or self.symb.analysis.n)

    def enquire(self):
        """(piv_order, d) in the format of NumericSubtree.hxx:424-470; for
        posdef returns (None, diag(L))."""
        lib = _lib.load()
        n = self.symb.analysis.n
        if self.posdef:
            d = np.zeros(n)
            lib.spral_ssids_gpu_subtree_enquire_dbl(True, self._h, None, _ptr(d))
            return None, d
        piv = np.zeros(n, dtype=np.int32)
        d = np.zeros(2 * n)
        lib.spral_ssids_gpu_subtree_enquire_dbl(False, self._h, _ptr(piv), _ptr(d))
        return piv, d

    def alter(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        _lib.load().spral_ssids_gpu_subtree_alter_dbl(self.posdef, self._h, _ptr(d))

    def get_contrib(self, device_resident=False):
        """contrib_type for the parent part (gpu/subtree.f90:522-538)."""
        c = Contrib()
        _lib.load().spral_ssids_b200_contrib_fill(C.byref(c), self.posdef, self._h, device_resident)
        return c

    def timings(self):
        ms = np.zeros(24)
        _lib.load().spral_ssids_gpu_subtree_get_timings(self._h, ms.ctypes.data_as(C.POINTER(C.c_double)), 24)
        return ms

    def close(self):
        if self._h:
            _lib.load().spral_ssids_gpu_destroy_num_subtree_dbl(self.posdef, self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------
# spral_ssids-level driver for one process (all parts on the visible GPUs)
# --------------------------------------------------------------------------

class Akeep:
    def __init__(self, analysis, subtrees, scaling=None):
        self.analysis, self.subtrees = analysis, subtrees
        self.scaling = scaling        # akeep%scaling: saved by the matching-based ordering (options%ordering = 2)


def analyse(n, ptr, row, order=None, nemin=32, ngpu=1, devices=None, options=None, val=None, ordering=None, **kw):
    """ssids_analyse (src/ssids/ssids.f90:148-389): ordering, symbolic
    factorisation, subtree partition, one SymbolicSubtree per part.
    ordering: None / 1 = METIS (or the user's `order`), 2 / "matching" = matching-based ordering
    (match_order_metis, needs `val`; its scaling is kept for factor(scaling="matching"))."""
    saved = None
    if ordering in (2, "matching"):
        if val is None:
            raise ValueError("the matching-based ordering needs the matrix values (SSIDS_ERROR_VAL)")
        from . import scaling as S
        order, saved, _ = S.match_order_metis(n, ptr, row, val)
    a = Analysis(n, ptr, row, order=order, nemin=nemin, ngpu=ngpu, **kw)
    devices = devices or list(range(ngpu))
    subtrees = []
    for p in range(a.nparts):
        loc = int(a.exec_loc[p])
        dev = devices[0] if loc <= 1 else devices[(loc - 2) % len(devices)]
        subtrees.append(SymbolicSubtree(a, p, device=dev, options=options))
    return Akeep(a, subtrees, saved)


def free_contrib(c):
    """contrib_free (src/ssids/contrib_free.f90:17-31): owner 1 = this engine."""
    if c is None or not c.owner_ptr:
        return
    if c.owner == 1:
        _lib.load().spral_ssids_gpu_subtree_free_contrib_dbl(bool(c.posdef), c.owner_ptr)
    elif getattr(c, "_free_hook", None):
        c._free_hook(c)


def compute_scaling(analysis, val, method, options=None):
    """options%scaling of ssids_factor (src/ssids/ssids.f90:899-1028): "hungarian" (= 1, MC64-type
    matching, scale_if_singular = options%action) or "equilib" (= 4, infinity-norm equilibration).
    Returns the scaling vector in the user's variable order."""
    from . import scaling as S
    a = analysis
    if method in ("hungarian", "mc64", 1):
        action = True if options is None else bool(options.action)
        s, _, flag, _ = S.hungarian_scale_sym(a.n, a.ptr, a.row, val, scale_if_singular=action)
        if flag == S.ERROR_SINGULAR:
            raise ValueError("matrix is structurally singular and options.action is false (SSIDS_ERROR_SINGULAR)")
        return s
    if method in ("auction", 2):
        return S.auction_scale_sym(a.n, a.ptr, a.row, val)[0]
    if method in ("equilib", "mc77", 4):
        return S.equilib_scale_sym(a.n, a.ptr, a.row, val)[0]
    raise ValueError(f"unknown scaling {method!r}")


class Fkeep:
    def __init__(self, akeep, posdef, numeric, inform, scaling):
        self.akeep, self.posdef, self.numeric, self.inform, self.scaling = akeep, posdef, numeric, inform, scaling


def factor(akeep, posdef, val, options=None, scaling=None, device_contrib=True):
    """ssids_factor -> fkeep%inner_factor (src/ssids/fkeep.F90:61-232): factor
    every part in order, handing contribution blocks child part -> parent part."""
    a = akeep.analysis
    sc = None
    if isinstance(scaling, str) and scaling in ("matching", "saved"):      # options%scaling = 3
        if akeep.scaling is None:
            raise ValueError("no scaling saved by analyse (SSIDS_ERROR_NO_SAVED_SCALING)")
        scaling = akeep.scaling
    elif isinstance(scaling, str):
        scaling = compute_scaling(a, val, scaling, options)
    if scaling is not None:   # fkeep%scaling(i) = scale(invp(i))  (ssids.f90:921-926)
        sc = np.ascontiguousarray(np.asarray(scaling, dtype=np.float64)[a.invp - 1])
    nparts = a.nparts
    slots = [None] * (nparts + 1)
    numeric = []
    inform = dict(flag=0, num_delay=0, num_factor=0, num_flops=0, num_neg=0, num_two=0,
                  maxfront=0, maxsupernode=0, matrix_rank=int(a.sptr[a.nnodes]) - 1,
                  not_first_pass=0, not_second_pass=0, cuda_error=0)
    for p in range(nparts):
        lo, hi = int(a.contrib_ptr[p]) - 1, int(a.contrib_ptr[p + 1]) - 1
        cc = [slots[i] for i in range(lo, hi)]
        ns = akeep.subtrees[p].factor(posdef, val, cc, options, sc)
        numeric.append(ns)
        # the consumer releases the contribution blocks it was handed
        # (spral_ssids_contrib_free_dbl, src/ssids/contrib_free.f90:17-48)
        for c in cc:
            free_contrib(c)
        st = ns.stats
        # cpu_copy_stats_out (src/ssids/cpu/cpu_iface.f90:74-94)
        if st.flag < 0:
            inform["flag"] = min(inform["flag"], st.flag)
            inform["cuda_error"] = st.cuda_error
            break
        inform["flag"] = max(inform["flag"], st.flag)
        for k in ("num_delay", "num_factor", "num_flops", "num_neg", "num_two",
                  "not_first_pass", "not_second_pass"):
            inform[k] += getattr(st, k)
        inform["maxfront"] = max(inform["maxfront"], st.maxfront)
        inform["maxsupernode"] = max(inform["maxsupernode"], st.maxsupernode)
        inform["matrix_rank"] -= st.num_zero
        idx = int(a.contrib_idx[p]) - 1
        if idx < nparts:
            c = ns.get_contrib(device_resident=device_contrib)
            c.ready = 1
            slots[idx] = c
    return Fkeep(akeep, posdef, numeric, inform, sc)


def _sym_matvec(a, val, X):
    """A X for the lower-triangular CSC pattern of the analysis (1-based ptr / row) and the values given to factor."""
    n = a.n
    ptr = np.asarray(a.ptr, dtype=np.int64) - 1
    row = np.asarray(a.row, dtype=np.int64) - 1
    col = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr))
    v = np.asarray(val, dtype=np.float64)
    Y = np.zeros_like(X)
    np.add.at(Y, row, v[:, None] * X[col, :])
    off = row != col
    np.add.at(Y, col[off], v[off][:, None] * X[row[off], :])
    return Y


def solve(fkeep, x, job=0, refine=0, val=None):
    """ssids_solve -> inner_solve_cpu (src/ssids/fkeep.F90:234-323).
    x: (n,) or (n, nrhs) Fortran-ordered; returns the solution (same shape).
    refine = k (job 0 only; needs `val`, the values that were factorised): k steps of iterative refinement with the
    same factors, x += solve(b - A x).  With threshold pivoting (u = 0.01) the scaled residual of the reference's CPU
    engine and of this one is 1e-13 .. 1e-12 on the shifted stencil matrices; one step brings both below 1e-15
    (DESIGN.md 5).  The reference leaves refinement to the caller (its driver does not refine either)."""
    a = fkeep.akeep.analysis
    x = np.asarray(x, dtype=np.float64)
    if a.n == 0:                      # trivial matrix: nothing to do (src/ssids/ssids.f90:1193)
        return x.copy()
    if refine and job == 0:
        if val is None:
            raise ValueError("solve(refine=k) needs val, the matrix values that were factorised")
        B = np.asfortranarray(x.reshape(a.n, -1))
        Xs = solve(fkeep, B, 0)
        for _ in range(int(refine)):
            Xs = Xs + solve(fkeep, np.asfortranarray(B - _sym_matvec(a, val, Xs)), 0)
        return Xs[:, 0] if x.ndim == 1 else Xs
    one = x.ndim == 1
    X = np.asfortranarray(x.reshape(a.n, -1))
    nrhs = X.shape[1]
    x2 = np.asfortranarray(X[a.invp - 1, :])
    if fkeep.scaling is not None and job in (0, 1):
        x2 *= fkeep.scaling[:, None]
    parts = fkeep.numeric
    if job in (0, 1):
        for ns in parts:
            ns.solve_fwd(x2, nrhs, a.n)
    if job == 2:
        for ns in parts:
            ns.solve_diag(x2, nrhs, a.n)
    if job == 3:
        for ns in reversed(parts):
            ns.solve_bwd(x2, nrhs, a.n)
    if job in (0, 4):
        for ns in reversed(parts):
            ns.solve_diag_bwd(x2, nrhs, a.n)
    if fkeep.scaling is not None and job in (0, 3, 4):
        x2 *= fkeep.scaling[:, None]
    out = np.empty_like(X)
    out[a.invp - 1, :] = x2
    return out[:, 0] if one else out
