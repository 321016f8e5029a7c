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


def _symbolic_cholesky(P):
    """Brute-force structure of the Cholesky factor of a symmetric 0/1 pattern (dense, right-looking)."""
    n = P.shape[0]
    S = P.copy().astype(bool)
    S[np.arange(n), np.arange(n)] = True
    for k in range(n):
        below = np.nonzero(S[k + 1:, k])[0] + k + 1
        for j in below:
            S[below[below >= j], j] = True
    return np.tril(S)


def test_row_lists_equal_the_true_fill_pattern():
    """Independent check of the restated basic_analyse (core_analyse.f90:38-156: etree, column counts,
    supernodes, row lists): with nemin = 1 only no-fill merges happen, so the row list of a supernode is
    EXACTLY the structure of L in its first column and num_factor is nnz(L); with nemin = 8 the row
    lists contain that structure.  The truth comes from a dense brute-force symbolic factorisation of
    the permuted pattern."""
    rng = np.random.default_rng(7)
    for trial in range(40):
        n = int(rng.integers(2, 45))
        R = sp.random(n, n, density=rng.uniform(0.02, 0.25), random_state=rng)
        A = sp.tril(R + R.T + sp.eye(n)).tocsc()
        A.sort_indices()
        ptr, row = A.indptr.astype(np.int64) + 1, A.indices.astype(np.int32) + 1
        order = (rng.permutation(n) + 1).astype(np.int32)            # any ordering, not only METIS
        for nemin in (1, 8):
            a = Analysis(n, ptr, row, order=order, nemin=nemin)
            pos = np.empty(n, dtype=np.int64)                        # pivot position (0-based) of each variable
            pos[a.invp - 1] = np.arange(n)
            F = (A + A.T).toarray() != 0
            Pm = np.zeros((n, n), dtype=bool)
            Pm[np.ix_(pos, pos)] = F
            L = _symbolic_cholesky(Pm)
            nn = a.nnodes
            for i in range(nn):
                c0, c1 = int(a.sptr[i]) - 1, int(a.sptr[i + 1]) - 1  # 0-based pivot positions [c0, c1)
                rows = a.rlist[a.rptr[i] - 1:a.rptr[i + 1] - 1] - 1
                for j in range(c0, c1):
                    truth = set(np.nonzero(L[:, j])[0].tolist())
                    mine = set(r for r in rows.tolist() if r >= j)
                    if nemin == 1:
                        assert truth == mine, (trial, n, i, j)
                    else:
                        assert truth <= mine, (trial, n, i, j)
            if nemin == 1:
                assert a.num_factor == int(L.sum())
            a.close()


def test_a_to_l_map_places_every_entry_where_it_belongs():
    """build_map (src/ssids/anal.F90:1137-1239): entry k of A, at (r, c) in the user's numbering, must land
    in the node that owns pivot column min(pos r, pos c), at local column = that pivot column and at the
    row of the node's row list that holds max(pos r, pos c)."""
    rng = np.random.default_rng(11)
    for trial in range(30):
        n = int(rng.integers(2, 70))
        R = sp.random(n, n, density=rng.uniform(0.02, 0.2), random_state=rng)
        A = sp.tril(R + R.T + sp.eye(n)).tocsc()
        A.sort_indices()
        ptr, row = A.indptr.astype(np.int64) + 1, A.indices.astype(np.int32) + 1
        for nemin in (1, 16):
            a = Analysis(n, ptr, row, nemin=nemin)
            pos = np.empty(n + 1, dtype=np.int64)
            pos[a.invp] = np.arange(1, n + 1)                        # 1-based pivot position of variable v
            col_of_entry = np.repeat(np.arange(1, n + 1), np.diff(ptr))
            nl = a.nlist.reshape(-1, 2)
            for i in range(a.nnodes):
                m = int(a.rptr[i + 1] - a.rptr[i])
                rows = a.rlist[a.rptr[i] - 1:a.rptr[i + 1] - 1]
                for src, dest in nl[a.nptr[i] - 1:a.nptr[i + 1] - 1]:
                    lc, lr = (dest - 1) // m, (dest - 1) % m          # local column / row, 0-based
                    r, c = int(row[src - 1]), int(col_of_entry[src - 1])
                    pr, pc = int(pos[r]), int(pos[c])
                    assert int(a.sptr[i]) + lc == min(pr, pc), (trial, i, src)
                    assert int(rows[lr]) == max(pr, pc), (trial, i, src)
            a.close()


def test_index_map_oracle_satisfies_the_definitions():
    """oracle/index_maps.py (what the device-built maps are compared with bit for bit on the GPU) against
    the definitions: rlist_direct(ii) is the position of the child's row in the parent's row list
    (gpu/subtree.f90:204-234); level(node) = num_levels - depth(node), roots last, deepest nodes in level 1,
    nodes of a level in increasing order (gpu/factor.f90:824-879)."""
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import index_maps
    from spral_b200 import matrices as M
    for gen in (lambda: M.stencil_3d_27pt(7, shift=13.0), lambda: M.laplacian_2d_5pt(23), lambda: M.kkt_saddle(600)):
        n, ptr, row, val = gen()
        a = Analysis(n, ptr, row, nemin=4)
        nn, spar, rp, rl, ncol = index_maps.part_view(a.sptr, a.sparent, a.rptr, a.rlist, 1, a.nnodes + 1)
        rd = index_maps.build_rlist_direct(n, nn, spar, rp, rl, ncol)
        for node in range(1, nn + 1):
            par = int(spar[node - 1])
            for ii in range(int(rp[node - 1]) + int(ncol[node - 1]), int(rp[node])):
                if par > nn:
                    assert rd[ii - 1] == -1
                else:
                    assert rl[int(rp[par - 1]) + int(rd[ii - 1]) - 2] == rl[ii - 1]
        nlev, lptr, llist = index_maps.assign_nodes_to_levels(nn, spar)
        depth = np.zeros(nn + 2, dtype=np.int64)
        for node in range(nn, 0, -1):                                 # parents have larger indices
            par = min(int(spar[node - 1]), nn + 1)
            depth[node] = 0 if par == nn + 1 else depth[par] + 1
        assert nlev == depth[1:nn + 1].max() + 1
        assert sorted(llist.tolist()) == list(range(1, nn + 1)) and lptr[0] == 1 and lptr[nlev] == nn + 1
        for lev in range(1, nlev + 1):
            nodes = llist[lptr[lev - 1] - 1:lptr[lev] - 1]
            assert np.all(np.diff(nodes) > 0)
            assert np.all(depth[nodes] == nlev - lev)
        cptr, clist = index_maps.build_child_pointers(nn, spar)
        for node in range(1, nn + 1):
            kids = clist[cptr[node - 1] - 1:cptr[node] - 1]
            assert np.all(np.diff(kids) > 0) and all(int(spar[k - 1]) == node for k in kids)
        a.close()
