"""CPU-only: pins the oracle (the reference's own CPU engine, oracle/_ref) against
the artefacts the reference holds for this path, and the restated analyse
against structural invariants."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_ref
from conftest import ROOT
from spral_b200 import matrices as M
from spral_b200.ssids import Analysis

pytestmark = pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")
GOLDEN = os.path.join(ROOT, "tests", "golden", "oracle_stats.json")


def factor_solve(gen, posdef, **kw):
    n, ptr, row, val = gen()
    a = Analysis(n, ptr, row, **kw)
    A = M.to_scipy(n, ptr, row, val)
    b = A @ np.arange(1.0, n + 1.0)
    parts, inform, sc = oracle_ref.ref_factor(a, posdef, val)
    x = oracle_ref.ref_solve(a, parts, posdef, b)
    for p in parts:
        p.close()
    return a, A, b, x, inform


def test_reference_kernel_unit_tests_pass():
    """The reference's own kernelst_cpp (tests/ssids/kernels.cxx) built against the
    same objects and BLAS the oracle library uses."""
    exe = os.path.join(ROOT, "oracle", "_ref", "kernelst")
    if not os.path.exists(exe):
        pytest.skip("kernelst not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "All tests passed" in out.stdout


def test_example_5x5_known_answer():
    """examples/C/ssids.c:18-31,54-56: the solution is (1,2,3,4,5); the documented
    run reports one negative eigenvalue and two 2x2 pivots."""
    n, ptr, row, val = M.example_5x5()
    a = Analysis(n, ptr, row)
    b = np.array([4.0, 17.0, 19.0, 2.0, 12.0])      # rhs of examples/C/ssids.c:29
    parts, inform, sc = oracle_ref.ref_factor(a, False, val)
    x = oracle_ref.ref_solve(a, parts, False, b)
    np.testing.assert_allclose(x, [1, 2, 3, 4, 5], rtol=0, atol=1e-13)
    assert (inform["num_neg"], inform["num_two"], inform["num_delay"]) == (1, 2, 0)
    assert (inform["num_factor"], inform["num_flops"]) == (15, 55)


@pytest.mark.parametrize("name,gen,posdef", [
    ("lap2d_100", lambda: M.laplacian_2d_5pt(100), True),          # BASELINE config 1
    ("lap3d_12", lambda: M.laplacian_3d_7pt(12), True),
    ("st27_12_s13", lambda: M.stencil_3d_27pt(12, shift=13.0), False),
    ("kkt_1500", lambda: M.kkt_saddle(1500), False),
])
def test_oracle_residual_meets_reference_tolerance(name, gen, posdef):
    """tests/ssids/ssids.f90:28,1685-1691: scaled residual < 5e-11."""
    a, A, b, x, inform = factor_solve(gen, posdef)
    assert inform["flag"] >= 0
    assert oracle_ref.backward_error(A, x, b) < 5e-11
    assert inform["matrix_rank"] == a.n


def test_oracle_matches_committed_golden_stats():
    """Golden vectors generated in this container by tests/golden/make_golden.py
    from the compiled reference; pins analyse + oracle against drift."""
    gold = json.load(open(GOLDEN))
    import make_golden
    for name, exp in gold.items():
        got = make_golden.run_case(name)
        for k in ("n", "nnodes", "num_factor", "num_flops", "maxfront", "num_neg", "matrix_rank"):
            assert got[k] == exp[k], (name, k, got[k], exp[k])


def test_analyse_invariants():
    for gen in (lambda: M.laplacian_2d_5pt(30), lambda: M.stencil_3d_27pt(9), lambda: M.kkt_saddle(800)):
        n, ptr, row, val = gen()
        a = Analysis(n, ptr, row)
        nn = a.nnodes
        sptr, sparent, rptr, rlist = a.sptr, a.sparent, a.rptr, a.rlist
        assert sptr[0] == 1 and sptr[nn] == n + 1
        assert all(sparent[i] > i + 1 for i in range(nn))              # postorder: parent after child
        nfact = nflops = 0
        for i in range(nn):
            rows = rlist[rptr[i] - 1:rptr[i + 1] - 1]
            ncol = sptr[i + 1] - sptr[i]
            assert list(rows[:ncol]) == list(range(sptr[i], sptr[i + 1]))    # own columns first
            assert all(np.diff(rows) > 0)                                     # elimination order
            m = len(rows)
            nfact += ncol * (ncol + 1) // 2 + ncol * (m - ncol)
            nflops += sum((m - j) ** 2 for j in range(ncol))
            if sparent[i] <= nn:                                              # contribution rows live in the parent
                prow = set(rlist[rptr[sparent[i] - 1] - 1:rptr[sparent[i]] - 1])
                assert set(rows[ncol:]) <= prow
        assert nfact == a.num_factor and nflops == a.num_flops                # calc_stats, core_analyse.f90:862-903
        # every A entry maps inside its node (build_map, anal.F90:1137-1239)
        nl = a.nlist.reshape(-1, 2)
        assert len(nl) == ptr[n] - 1
        assert sorted(nl[:, 0]) == list(range(1, len(nl) + 1))
        for i in range(nn):
            m = rptr[i + 1] - rptr[i]
            ncol = sptr[i + 1] - sptr[i]
            d = nl[a.nptr[i] - 1:a.nptr[i + 1] - 1, 1]
            assert d.min(initial=1) >= 1 and d.max(initial=1) <= m * ncol
        a.close()


def test_partition_covers_tree():
    n, ptr, row, val = M.laplacian_3d_7pt(16)
    for ngpu in (1, 2, 4, 8):
        a = Analysis(n, ptr, row, ngpu=ngpu)
        assert a.part[0] == 1 and a.part[a.nparts] == a.nnodes + 1
        assert all(np.diff(a.part) > 0)
        # every part hands its contribution to a LATER part (or to nobody)
        for p in range(a.nparts):
            idx = a.contrib_idx[p]
            assert idx == a.nparts + 1 or idx >= 1
            root = a.part[p + 1] - 1
            par = a.sparent[root - 1]
            if par <= a.nnodes:
                q = np.searchsorted(a.part, par, side="right") - 1
                assert q > p
                assert a.contrib_dest[idx - 1] == par
        a.close()
