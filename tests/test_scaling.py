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


def test_spral_random_lcg_known_answers():
    """The reference's generator is the ANSI C / glibc TYPE_0 LCG x <- (1103515245 x + 12345) mod 2^31
    (src/random.f90:19-22); its sequence from seed 1 is the published one."""
    r = M.SpralRandom(1)
    seq = []
    for _ in range(5):
        r.integer(10)
        seq.append(r.state)
    assert seq == [1103527590, 377401575, 662824084, 1147902781, 2035015474]
    r = M.SpralRandom()                       # default seed 486502
    x1 = (1103515245 * 486502 + 12345) % 2 ** 31
    assert r.real() == 1.0 - 2.0 * x1 / 2.0 ** 31
    assert 1 <= M.SpralRandom(7).integer(13) <= 13


@pytest.mark.parametrize("mtype,m,n,nnz", [(M.MATRIX_REAL_SYM_INDEF, 50, 50, 220), (M.MATRIX_REAL_SYM_PSDEF, 9, 9, 45),
                                           (M.MATRIX_REAL_UNSYM, 30, 30, 100), (M.MATRIX_REAL_RECT, 12, 40, 90)])
def test_random_matrix_generate_invariants(mtype, m, n, nnz):
    """What tests/random_matrix.f90 checks of the reference's generator: entry count, row range,
    no duplicates, sorted columns, the diagonal of a symmetric non-singular pattern."""
    st = M.SpralRandom()
    ptr, row, val = M.random_matrix_generate(st, mtype, m, n, nnz, nonsingular=True, sort=True)
    sym = mtype in (M.MATRIX_REAL_SYM_INDEF, M.MATRIX_REAL_SYM_PSDEF)
    assert ptr[0] == 1 and ptr[n] - 1 == nnz == len(row) == len(val)
    assert np.all(np.abs(val) <= 1.0)
    for j in range(n):
        col = row[ptr[j] - 1:ptr[j + 1] - 1]
        assert list(col) == sorted(set(col)) and (len(col) == 0 or (col[0] >= (j + 1 if sym else 1) and col[-1] <= m))
        if sym:
            assert col[0] == j + 1                        # forced diagonal comes first after sorting
    if not sym:                                           # structurally non-singular: a full matching exists
        A = sp.csc_matrix((np.ones(nnz), row - 1, ptr - 1), shape=(m, n))
        from scipy.sparse.csgraph import maximum_bipartite_matching
        assert (maximum_bipartite_matching(A.tocsr(), perm_type="column") >= 0).sum() == min(m, n)
    # same state -> same matrix
    ptr2, row2, val2 = M.random_matrix_generate(M.SpralRandom(), mtype, m, n, nnz, nonsingular=True, sort=True)
    assert np.array_equal(ptr, ptr2) and np.array_equal(row, row2) and np.array_equal(val, val2)


def test_reference_test_generators():
    st = M.SpralRandom()
    n, ptr, row, val = M.gen_random_posdef(st, 40, 160)
    A = M.to_scipy(n, ptr, row, val).toarray()
    assert np.all(np.linalg.eigvalsh(A) > 0)              # diagonally dominant
    n, ptr, row, val = M.gen_random_indef(st, 40, 160)
    A = M.to_scipy(n, ptr, row, val).toarray()
    assert (np.diag(A) == 0).any() and np.abs(A).max() > 10.0


def test_matching_based_ordering_pairs_are_adjacent_and_remove_delays():
    """match_order_metis (src/match_order.f90:51-396): a permutation in which the two variables of every
    matched 2-cycle are consecutive; with its scaling the reference CPU engine factorises KKT matrices
    without a single delayed pivot (METIS alone: hundreds)."""
    import oracle_ref
    from spral_b200.ssids import Analysis
    oracle_ref.ensure_env()
    n, ptr, row, val = M.kkt_grid(12)
    m = n - 12 ** 3
    order, s, flag = S.match_order_metis(n, ptr, row, val)
    assert flag == 0 and sorted(order.tolist()) == list(range(1, n + 1))
    _, match, _, _ = S.hungarian_scale_sym(n, ptr, row, val)
    pairs = [(i, match[i] - 1) for i in range(n) if match[match[i] - 1] - 1 == i and match[i] - 1 != i]
    assert pairs and all(abs(int(order[i]) - int(order[j])) == 1 for i, j in pairs)
    d = M.to_scipy(n, ptr, row, val).diagonal()
    assert all(match[i] - 1 != i for i in range(n) if d[i] == 0.0)
    delays = {}
    for label, kw, sc in (("metis", {}, None), ("matching", {"order": order}, s)):
        a = Analysis(n, ptr, row, **kw)
        parts, r, _ = oracle_ref.ref_factor(a, False, val, scaling=sc)
        for p in parts:
            p.close()
        assert r["flag"] == 0 and r["num_neg"] == m and r["matrix_rank"] == n
        delays[label] = r["num_delay"]
        a.close()
    assert delays["matching"] == 0 and delays["metis"] > 20, delays


def test_matching_based_ordering_on_singular_and_definite_matrices():
    n, ptr, row, val = M.laplacian_2d_5pt(9)                  # positive definite: identity matching
    order, s, flag = S.match_order_metis(n, ptr, row, val)
    assert flag == 0 and sorted(order.tolist()) == list(range(1, n + 1))
    rng = np.random.default_rng(12)
    A = _random_sym(30, 0.2, rng).tolil()
    A[4, :] = 0.0
    A[:, 4] = 0.0
    A = A.tocsc(); A.eliminate_zeros()
    n, ptr, row, val = _lower_csc(A)
    order, s, flag = S.match_order_metis(n, ptr, row, val)
    assert flag == 1 and sorted(order.tolist()) == list(range(1, n + 1)) and np.all(np.isfinite(s))


@pytest.mark.parametrize("n,density,seed", [(40, 0.15, 0), (300, 0.02, 1), (2000, 0.003, 2)])
def test_auction_scaling(n, density, seed):
    """auction_scale_sym (src/scaling.f90:269-309): an approximate matching (the reference's tests accept
    >= 90 % matched, tests/scaling.f90) whose prices scale the entries to O(1)."""
    rng = np.random.default_rng(seed)
    A = _random_sym(n, density, rng)
    nn, ptr, row, val = _lower_csc(A)
    s, match, matched, it = S.auction_scale_sym(nn, ptr, row, val)
    assert matched >= 0.9 * n and it >= 1
    m = match[match > 0]
    assert len(set(m.tolist())) == len(m) == matched             # a (partial) matching
    assert np.all(s > 0) and np.all(np.isfinite(s))
    B = abs(sp.diags(s) @ A @ sp.diags(s)).tocsc()
    # epsilon-optimal duals: scaled entries are bounded by exp(eps) with eps <= 1
    assert B.max() <= np.e * (1 + 1e-12)
    s_h, _, _, _ = S.hungarian_scale_sym(nn, ptr, row, val)
    Bh = abs(sp.diags(s_h) @ A @ sp.diags(s_h))
    assert B.max() < 1e3 * Bh.max()


def test_fuzz_orderings_and_scalings_on_random_matrices():
    """Random symmetric patterns with values over 12 orders of magnitude, missing diagonals, structurally
    singular ones included: every routine returns finite positive scalings, match_order_metis a permutation,
    and on non-singular matrices the matching scaling keeps |S A S| <= 1."""
    rng = np.random.default_rng(5)
    done = singular = 0
    for trial in range(150):
        n = int(rng.integers(1, 80))
        R = sp.random(n, n, density=rng.uniform(0.01, 0.3), random_state=rng,
                      data_rvs=lambda k: rng.uniform(-1, 1, k) * 10.0 ** rng.integers(-6, 7, k))
        A = (R + R.T).tolil()
        for i in range(n):
            if rng.uniform() < 0.6:
                A[i, i] = rng.uniform(-2, 2)
        A = A.tocsc()
        A.eliminate_zeros()
        if A.nnz == 0:
            continue
        nn, ptr, row, val = _lower_csc(A)
        order, s, flag = S.match_order_metis(nn, ptr, row, val)
        assert sorted(order.tolist()) == list(range(1, n + 1))
        assert np.all(np.isfinite(s)) and np.all(s > 0)
        if flag == 0:
            assert abs(sp.diags(s) @ A @ sp.diags(s)).max() <= 1 + 1e-9
        else:
            singular += 1
        for sc in (S.auction_scale_sym(nn, ptr, row, val)[0], S.equilib_scale_sym(nn, ptr, row, val)[0]):
            assert np.all(np.isfinite(sc)) and np.all(sc > 0)
        done += 1
    assert done > 100 and singular > 5
