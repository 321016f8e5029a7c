"""CPU tests of the host scaling pre-processing (spral_b200/scaling.py, csrc/scaling.cpp),
modelled on the reference's tests/scaling.f90 (random matrices; scaled entries <= 1, the
matched entries == 1, singular matrices flagged) plus an optimality check of the matching
against scipy's dense linear_sum_assignment.  Parity with the reference's own duals is
unpinned (no Fortran compiler; the optimal duals are not unique)."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linear_sum_assignment

from spral_b200 import matrices as M
from spral_b200 import scaling as S


def _lower_csc(A):
    A = sp.tril(sp.csc_matrix(A)).tocsc()
    A.sort_indices()
    return A.shape[0], A.indptr.astype(np.int64) + 1, A.indices.astype(np.int32) + 1, A.data.astype(np.float64)


def _random_sym(n, density, rng, zero_diag_from=None):
    R = sp.random(n, n, density=density, random_state=rng, data_rvs=lambda k: rng.uniform(-1, 1, k) * 10.0 ** rng.integers(-4, 5, k))
    A = (R + R.T).tolil()
    for i in range(n):
        A[i, i] = 0.0 if (zero_diag_from is not None and i >= zero_diag_from) else rng.uniform(1, 2) * 10.0 ** rng.integers(-3, 4)
    A = A.tocsc()
    A.eliminate_zeros()
    return A


@pytest.mark.parametrize("n,density,seed", [(30, 0.2, 0), (60, 0.1, 1), (200, 0.03, 2), (1000, 0.005, 3)])
def test_hungarian_scaling_property_and_optimality(n, density, seed):
    rng = np.random.default_rng(seed)
    A = _random_sym(n, density, rng)
    nn, ptr, row, val = _lower_csc(A)
    s, match, flag, matched = S.hungarian_scale_sym(nn, ptr, row, val)
    assert flag == 0 and matched == n
    assert sorted(match) == list(range(1, n + 1))              # a permutation
    B = abs(sp.diags(s) @ A @ sp.diags(s)).tocsc()
    assert B.max() <= 1.0 + 1e-10
    # the matching maximises the product of the matched entries: compare with the dense optimum
    if n <= 200:
        D = abs(A).toarray()
        with np.errstate(divide="ignore"):
            cost = np.where(D > 0, -np.log(D), 1e6)
        r, c = linear_sum_assignment(cost)
        best = cost[r, c].sum()
        mine = sum(cost[i, match[i] - 1] for i in range(n))
        assert mine <= best + 1e-8 * max(1.0, abs(best))
    # unsymmetric view of the same duals: every row and column of the scaled matrix reaches 1
    # on its matched entry up to the symmetrisation; check the weaker, exact statement:
    assert np.all(B.max(axis=0).toarray().ravel() > 0)


def test_hungarian_on_kkt_matrix_and_factor_input():
    n, ptr, row, val = M.kkt_saddle(400, 0.3, seed=5)
    s, match, flag, matched = S.hungarian_scale_sym(n, ptr, row, val)
    assert flag == 0 and matched == n
    A = M.to_scipy(n, ptr, row, val)
    B = abs(sp.diags(s) @ A @ sp.diags(s))
    assert B.max() <= 1.0 + 1e-10
    # every constraint row (zero diagonal) is matched off the diagonal
    d = A.diagonal()
    assert all(match[i] - 1 != i for i in range(n) if d[i] == 0.0)


def test_hungarian_structurally_singular():
    rng = np.random.default_rng(9)
    A = _random_sym(40, 0.15, rng).tolil()
    for i in (7, 23):                                          # two empty rows / columns
        A[i, :] = 0.0
        A[:, i] = 0.0
    A = A.tocsc(); A.eliminate_zeros()
    n, ptr, row, val = _lower_csc(A)
    s, match, flag, matched = S.hungarian_scale_sym(n, ptr, row, val)
    assert flag == S.ERROR_SINGULAR and matched == n - 2 and np.all(s == 1.0)
    s, match, flag, matched = S.hungarian_scale_sym(n, ptr, row, val, scale_if_singular=True)
    assert flag == S.WARNING_SINGULAR and matched == n - 2
    assert match[7] < 0 and match[23] < 0 and np.all(np.isfinite(s)) and np.all(s > 0)
    assert s[7] == 1.0 and s[23] == 1.0                        # nothing to scale against: 1/0 -> 1
    B = abs(sp.diags(s) @ A @ sp.diags(s))
    assert B.max() <= 1.0 + 1e-10


def test_equilibration_converges_to_unit_inf_norms():
    rng = np.random.default_rng(4)
    A = _random_sym(300, 0.02, rng)
    n, ptr, row, val = _lower_csc(A)
    s, it = S.equilib_scale_sym(n, ptr, row, val, max_iterations=50, tol=1e-8)
    B = abs(sp.diags(s) @ A @ sp.diags(s)).tocsc()
    rmax = B.max(axis=0).toarray().ravel()
    assert it <= 50 and np.abs(rmax - 1).max() < 1e-3
    s10, it10 = S.equilib_scale_sym(n, ptr, row, val)          # the reference's defaults: 10 sweeps
    assert it10 <= 10 and np.all(s10 > 0)
