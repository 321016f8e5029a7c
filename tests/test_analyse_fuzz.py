"""CPU fuzz of the restated analyse phase (spral_b200/csrc/analyse.cpp: expand_pattern, basic_analyse,
build_map, find_subtree_partition) on small random patterns, including the degenerate ones the
reference's test-suite feeds ssids_analyse (tests/ssids/ssids.f90: missing diagonals, empty columns,
the empty matrix): no crash, and the structural invariants of akeep (src/ssids/akeep.f90:25-80)."""
import numpy as np
import scipy.sparse as sp

from spral_b200.ssids import Analysis


def test_random_patterns_keep_the_akeep_invariants():
    rng = np.random.default_rng(1)
    for trial in range(150):
        n = int(rng.integers(1, 60))
        R = sp.random(n, n, density=rng.uniform(0, 0.3), random_state=rng)
        A = sp.tril(R + R.T).tolil()
        keepdiag = rng.uniform() < 0.7
        for i in range(n):
            if keepdiag or rng.uniform() < 0.5:
                A[i, i] = 1.0
        A = A.tocsc()
        A.sort_indices()
        ptr, row = A.indptr.astype(np.int64) + 1, A.indices.astype(np.int32) + 1
        for nemin, ngpu in ((1, 1), (8, 4), (32, 1)):
            a = Analysis(n, ptr, row, nemin=nemin, ngpu=ngpu)
            nn = a.nnodes
            assert sorted(a.invp.tolist()) == list(range(1, n + 1))
            if nn:
                assert a.sptr[0] == 1 and np.all(np.diff(a.sptr) > 0)
                assert np.all(a.sparent > np.arange(1, nn + 1))            # parents after children (postorder)
                assert a.part[0] == 1 and a.part[a.nparts] == nn + 1
                src = a.nlist[0::2][: int(a.nptr[nn]) - 1]
                assert len(set(src.tolist())) == len(src)                  # every mapped entry is mapped once
                m = a.rptr[1:] - a.rptr[:-1]
                nc = a.sptr[1:] - a.sptr[:-1]
                assert np.all(m >= nc)
            a.close()


def test_matrices_without_entries():
    e = np.zeros(0, dtype=np.int32)
    a = Analysis(1, np.array([1, 1], dtype=np.int64), e)                  # 1 x 1, no entry
    assert a.nnodes == 0 and a.nparts == 0 and a.flag == 6                # SSIDS_WARNING_ANAL_SINGULAR
    a = Analysis(3, np.ones(4, dtype=np.int64), e)
    assert a.nnodes == 0 and a.invp.tolist() == [1, 2, 3]
    a = Analysis(0, np.array([1], dtype=np.int64), e)
    assert a.nnodes == 0 and a.flag == 0
