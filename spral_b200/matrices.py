"""Synthetic test matrices of BASELINE.json (lower-triangular CSC, 1-based).

Every generator returns (n, ptr[int64, n+1], row[int32], val[float64]) holding
the LOWER triangle column by column with 1-based index values, i.e. exactly the
arrays `ssids_analyse`/`ssids_factor` take (docs/Fortran/ssids.rst, `ptr,row,val`).
"""
import numpy as np
import scipy.sparse as sp


def _lower_csc(A):
    A = sp.tril(sp.csc_matrix(A), format="csc")
    A.sort_indices()
    n = A.shape[0]
    return (n, (A.indptr.astype(np.int64) + 1), (A.indices.astype(np.int32) + 1),
            A.data.astype(np.float64))


def to_scipy(n, ptr, row, val):
    """Full symmetric scipy CSC from the lower-triangular 1-based arrays."""
    L = sp.csc_matrix((val, row - 1, ptr - 1), shape=(n, n))
    return (L + sp.tril(L, -1).T).tocsc()


def laplacian_2d_5pt(nx, ny=None):
    """cfg1: 2-D 5-point Laplacian, diag 4, off-diagonals -1 (posdef)."""
    ny = ny or nx
    ex, ey = np.ones(nx), np.ones(ny)
    Tx = sp.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
    Ty = sp.diags([-ey[:-1], 2 * ey, -ey[:-1]], [-1, 0, 1])
    A = sp.kron(sp.eye(ny), Tx) + sp.kron(Ty, sp.eye(nx))
    return _lower_csc(A)


def laplacian_3d_7pt(nx, ny=None, nz=None):
    """cfg2: 3-D 7-point Laplacian, diag 6, off-diagonals -1 (posdef)."""
    ny = ny or nx
    nz = nz or nx

    def T(k):
        e = np.ones(k)
        return sp.diags([-e[:-1], 2 * e, -e[:-1]], [-1, 0, 1])
    A = (sp.kron(sp.eye(nz), sp.kron(sp.eye(ny), T(nx))) +
         sp.kron(sp.eye(nz), sp.kron(T(ny), sp.eye(nx))) +
         sp.kron(T(nz), sp.kron(sp.eye(ny), sp.eye(nx))))
    return _lower_csc(A)


def stencil_3d_27pt(nx, ny=None, nz=None, shift=0.0):
    """cfg3/cfg5: 3-D 27-point stencil, diag 26, all 26 neighbours -1, minus
    shift*I.  shift=0 is positive semi-definite-ish (diag dominant, posdef);
    a shift inside the spectrum (e.g. 13) makes it indefinite."""
    ny = ny or nx
    nz = nz or nx

    def B(k):   # tridiagonal of ones (incl. diagonal)
        e = np.ones(k)
        return sp.diags([e[:-1], e, e[:-1]], [-1, 0, 1])
    N = sp.kron(B(nz), sp.kron(B(ny), B(nx)))          # 27 ones incl. centre
    n = nx * ny * nz
    A = -N + (27.0 - shift) * sp.eye(n)                  # centre: -1 + 27 - shift = 26 - shift
    return _lower_csc(A)


class SpralRandom:
    """Restatement of the reference's LCG (src/random.f90:15-22,55-80):
    x <- (1103515245*x + 12345) mod 2^31, default seed 486502."""
    A, C_, M = 1103515245, 12345, 2 ** 31

    def __init__(self, seed=486502):
        self.state = seed

    def real(self, positive=False):
        self.state = (self.A * self.state + self.C_) % self.M
        if positive:
            return float(self.state) / float(self.M)
        return 1.0 - 2.0 * float(self.state) / float(self.M)

    def integer(self, n):
        self.state = (self.A * self.state + self.C_) % self.M
        return int(self.state * n // self.M) + 1


    def integer_in_range(self, lo, hi):
        """random_integer_in_range (src/random_matrix.f90:302-309)."""
        return lo + self.integer(hi - lo + 1) - 1

    def sym_wt_integer(self, n):
        """random_sym_wt_integer (src/random_matrix.f90:281-297): column index weighted by the
        number of entries of the column in the lower triangle."""
        r1, r2 = self.integer(n), self.integer(n)
        while r2 < r1:
            r1, r2 = self.integer(n), self.integer(n)
        return r1

    def perm(self, n):
        """random_perm (src/random_matrix.f90:314-335), 1-based values."""
        p = list(range(1, n + 1))
        for i in range(1, n):
            j = self.integer_in_range(i, n)
            p[i - 1], p[j - 1] = p[j - 1], p[i - 1]
        return p


MATRIX_REAL_RECT, MATRIX_REAL_UNSYM, MATRIX_REAL_SYM_PSDEF, MATRIX_REAL_SYM_INDEF = 1, 2, 3, 4


def random_matrix_generate(state, matrix_type, m, n, nnz, nonsingular=False, sort=False):
    """Restatement of random_matrix_generate (src/random_matrix.f90:84-279) on a SpralRandom
    state: random m x n CSC pattern with nnz entries (lower triangle for the symmetric types),
    optionally structurally non-singular and with sorted columns, values uniform in [-1, 1].
    Returns (ptr, row, val) with 1-based ptr / row as the reference does.  Pure Python: meant
    for the small problems of the reference's own tests (tests/ssids/ssids.f90: n <= 1000)."""
    sym = matrix_type in (MATRIX_REAL_SYM_PSDEF, MATRIX_REAL_SYM_INDEF)
    if sym and m != n:
        raise ValueError("symmetric matrix must be square")
    if m < 1 or n < 1 or nnz < 1 or (sym and n * (n + 1) // 2 < nnz) or (not sym and m * n < nnz):
        raise ValueError("arguments out of range")
    if nonsingular and nnz < min(m, n):
        raise ValueError("not enough entries for a non-singular matrix")
    cnt = [0] * (n + 1)                                   # 1-based
    rperm = cperm = None
    if sym:
        if nonsingular:
            rperm = list(range(1, m + 1))
            cperm = list(range(1, n + 1))
            for j in range(1, n + 1):
                cnt[j] += 1
        for _ in range(nnz - (min(m, n) if nonsingular else 0)):
            j = state.sym_wt_integer(n)
            while cnt[j] >= m - j + 1:
                j = state.sym_wt_integer(n)
            cnt[j] += 1
    else:
        if nonsingular:
            rperm, cperm = state.perm(m), state.perm(n)
            for j in range(1, n + 1):
                if cperm[j - 1] <= min(m, n):
                    cnt[j] = 1
        for _ in range(nnz - (min(m, n) if nonsingular else 0)):
            j = state.integer(n)
            while cnt[j] >= m:
                j = state.integer(n)
            cnt[j] += 1
    ptr = [1] * (n + 2)
    row = [0] * (nnz + 1)
    rused = [False] * (m + 1)
    for i in range(1, n + 1):
        ptr[i + 1] = ptr[i] + cnt[i]
        jj = ptr[i]
        if nonsingular and cperm[i - 1] <= min(m, n):
            k = rperm[cperm[i - 1] - 1]
            row[jj] = k
            rused[k] = True
            jj += 1
        minidx = i if sym else 1
        for q in range(jj, ptr[i + 1]):
            k = state.integer_in_range(minidx, m)
            while rused[k]:
                k = state.integer_in_range(minidx, m)
            row[q] = k
            rused[k] = True
        for q in range(ptr[i], ptr[i + 1]):
            rused[row[q]] = False
    if sort:                                              # dbl_tr_sort: increasing row order per column
        for i in range(1, n + 1):
            row[ptr[i]:ptr[i + 1]] = sorted(row[ptr[i]:ptr[i + 1]])
    val = [state.real() for _ in range(ptr[n + 1] - 1)]
    return (np.asarray(ptr[1:n + 2], dtype=np.int64), np.asarray(row[1:], dtype=np.int32),
            np.asarray(val, dtype=np.float64))


def gen_random_posdef(state, n, nza):
    """gen_random_posdef (tests/ssids/ssids.f90:2872-2897): sorted non-singular random pattern made
    diagonally dominant (the first entry of a sorted column is its diagonal)."""
    ptr, row, val = random_matrix_generate(state, MATRIX_REAL_SYM_PSDEF, n, n, nza, nonsingular=True, sort=True)
    for k in range(1, n + 1):
        tempv = 0.0
        for j in range(ptr[k - 1] + 1, ptr[k]):           # 1-based positions ptr(k)+1 .. ptr(k+1)-1
            tempv += abs(val[j - 1])
            i = ptr[row[j - 1] - 1]
            val[i - 1] += abs(val[j - 1])
        i = ptr[k - 1]
        val[i - 1] = 1.0 + val[i - 1] + tempv
    return n, ptr, row, val


def gen_random_indef(state, n, nza):
    """gen_random_indef (tests/ssids/ssids.f90:2829-2868) without the zero-row option: some
    explicit zeros on the diagonal and a large last off-diagonal entry in every column."""
    ptr, row, val = random_matrix_generate(state, MATRIX_REAL_SYM_INDEF, n, n, nza, nonsingular=True, sort=True)
    if n > 3:
        step = max(1, state.integer(n // 2))
        for k in range(1, n + 1, step):
            if ptr[k] > ptr[k - 1] + 1:
                val[ptr[k - 1] - 1] = 0.0
        for k in range(1, n + 1):
            val[ptr[k] - 2] *= 1000.0
    return n, ptr, row, val


def kkt_saddle(n, frac_constraints=0.3, nnz_per_row=6, seed=486502):
    """cfg4: synthetic KKT saddle-point matrix [H B^T; B 0] of order n with
    m = frac*n zero-diagonal constraint rows.  H is a sparse diagonally dominant
    SPD block, B a sparse full-row-rank block; entries are drawn with the
    restated spral_random LCG for scalars and numpy (seeded from it) for bulk
    patterns, so the matrix is deterministic."""
    m = int(round(frac_constraints * n))
    nh = n - m
    lcg = SpralRandom(seed)
    rng = np.random.default_rng(lcg.integer(2 ** 30))
    # H: random symmetric pattern + dominant diagonal
    k = nnz_per_row // 2
    r = np.repeat(np.arange(nh), k)
    c = rng.integers(0, nh, size=nh * k)
    v = rng.uniform(-1.0, 1.0, size=nh * k)
    Hoff = sp.coo_matrix((v, (r, c)), shape=(nh, nh)).tocsr()
    Hoff = sp.tril(Hoff, -1)
    Hoff = Hoff + Hoff.T
    d = np.asarray(abs(Hoff).sum(axis=1)).ravel() + 1.0
    H = Hoff + sp.diags(d)
    # B: each constraint couples nnz_per_row primal variables; a shifted identity
    # makes it full row rank
    rb = np.repeat(np.arange(m), nnz_per_row)
    cb = rng.integers(0, nh, size=m * nnz_per_row)
    vb = rng.uniform(-1.0, 1.0, size=m * nnz_per_row)
    Bm = sp.coo_matrix((vb, (rb, cb)), shape=(m, nh)).tocsr()
    Bm = Bm + sp.coo_matrix((np.full(m, 2.0), (np.arange(m), np.arange(m) % nh)), shape=(m, nh))
    K = sp.bmat([[H, Bm.T], [Bm, None]], format="csc")
    # explicit zero diagonal for the (2,2) block so that pivots are well defined
    K = K + sp.csc_matrix((np.zeros(m), (np.arange(nh, n), np.arange(nh, n))), shape=(n, n))
    return _lower_csc_keep_zeros(K)


def kkt_grid(g, frac_constraints=0.3, seed=486502):
    """cfg4 at benchmark size: a KKT saddle-point matrix [H B^T; B 0] with LOCAL coupling, so that it
    can be factorised at n ~ 500 000 (g = 70).  H is the 3-D 7-point Laplacian on a g^3 grid (SPD),
    each of the m = frac/(1-frac) g^3 constraints couples one grid node (drawn without repetition) with
    its +x, +y, +z neighbours (a perturbed discrete divergence; B has full row rank), the (2,2) block is
    an explicit zero diagonal.  Inertia is exactly (g^3 positive, m negative).  The random-pattern
    matrices of kkt_saddle / random_matrix_generate are expanders: their factors grow like n^2
    (measured: n = 100 000 -> 3.2e13 flops, 39 000-row front) and cannot be formed at n = 500 000."""
    nh = g ** 3
    m = int(round(frac_constraints / (1.0 - frac_constraints) * nh))
    n_, ptr, row, val = laplacian_3d_7pt(g)
    H = to_scipy(n_, ptr, row, val).tocsr()
    rng = np.random.default_rng(SpralRandom(seed).integer(2 ** 30))
    cells = np.sort(rng.choice(nh, size=m, replace=False))
    idx = np.arange(nh).reshape(g, g, g)
    nb = [np.roll(idx, -1, axis=ax).ravel() for ax in range(3)]
    rows = np.repeat(np.arange(m), 4)
    cols = np.stack([cells, nb[0][cells], nb[1][cells], nb[2][cells]], axis=1).ravel()
    vals = np.stack([np.full(m, 1.0)] + [rng.uniform(-0.6, -0.2, m) for _ in range(3)], axis=1).ravel()
    Bm = sp.coo_matrix((vals, (rows, cols)), shape=(m, nh)).tocsr()
    n = nh + m
    K = sp.bmat([[H, Bm.T], [Bm, None]], format="csc")
    K = K + sp.csc_matrix((np.zeros(m), (np.arange(nh, n), np.arange(nh, n))), shape=(n, n))
    return _lower_csc_keep_zeros(K)


def _lower_csc_keep_zeros(A):
    A = sp.csc_matrix(A)
    A.sort_indices()
    n = A.shape[0]
    ptr = [0]
    rows, vals = [], []
    indptr, indices, data = A.indptr, A.indices, A.data
    keep = np.zeros(len(indices), dtype=bool)
    col = np.repeat(np.arange(n), np.diff(indptr))
    keep = indices >= col
    newcounts = np.bincount(col[keep], minlength=n)
    ptr = np.concatenate([[0], np.cumsum(newcounts)]).astype(np.int64) + 1
    return n, ptr, (indices[keep].astype(np.int32) + 1), data[keep].astype(np.float64)


def example_5x5():
    """The matrix of examples/C/ssids.c:18-31 (solution of its rhs is 1..5)."""
    ptr = np.array([1, 3, 6, 8, 9, 10], dtype=np.int64)
    row = np.array([1, 2, 2, 3, 5, 3, 4, 4, 5], dtype=np.int32)
    val = np.array([2.0, 1.0, 4.0, 1.0, 1.0, 3.0, 2.0, -1.0, 2.0])
    return 5, ptr, row, val
