"""TEST INFRASTRUCTURE (oracle) -- numpy/pure-Python restatement of the
reference's GPU-side symbolic maps, used only by tests/ to check the engine's
device-built maps bit for bit.  Not imported by the product.

Follows (reference tree ralna/spral, 1-based values kept):
  build_child_pointers     src/ssids/gpu/subtree.f90:162-196
  build_rlist_direct       src/ssids/gpu/subtree.f90:204-234
  assign_nodes_to_levels   src/ssids/gpu/factor.f90:824-879

parity unpinned at this boundary: the reference's host code is Fortran and
cannot be executed in this environment (no Fortran compiler); this file is a
line-by-line restatement, checked by structural invariants in tests/.
"""
import numpy as np


def part_view(sptr, sparent, rptr, rlist, sa, en):
    """Slices of the analyse arrays for the part of nodes sa..en-1 (1-based),
    with parents outside the part mapped to nnodes+1 as the reference does
    (src/ssids/gpu/subtree.f90:115-122, min(sparent, nnodes+1))."""
    nn = en - sa
    sp = np.array([min(int(sparent[sa - 1 + i]) - sa + 1, nn + 1) if sparent[sa - 1 + i] < en else nn + 1
                   for i in range(nn)], dtype=np.int64)
    rp = np.asarray(rptr[sa - 1:en], dtype=np.int64) - int(rptr[sa - 1]) + 1
    rl = np.asarray(rlist[int(rptr[sa - 1]) - 1:int(rptr[en - 1]) - 1], dtype=np.int64)
    ncol = np.asarray(sptr[sa:en], dtype=np.int64) - np.asarray(sptr[sa - 1:en - 1], dtype=np.int64)
    return nn, sp, rp, rl, ncol


def build_child_pointers(nnodes, sparent):
    child_head = [-1] * (nnodes + 2)
    child_next = [0] * (nnodes + 2)
    for i in range(nnodes, 0, -1):          # backwards so child list is in order
        j = min(int(sparent[i - 1]), nnodes + 1)
        child_next[i] = child_head[j]
        child_head[j] = i
    child_ptr = [1]
    child_list = []
    for i in range(1, nnodes + 2):
        j = child_head[i]
        while j != -1:
            child_list.append(j)
            j = child_next[j]
        child_ptr.append(len(child_list) + 1)
    return np.array(child_ptr), np.array(child_list)


def build_rlist_direct(n, nnodes, sparent, rptr, rlist, ncol):
    """Returns rlist_direct with -1 in the positions the reference leaves
    undefined (the child's own columns, which are not rows of the parent)."""
    out = np.full(len(rlist), -1, dtype=np.int64)
    for node in range(1, nnodes + 1):
        parent = int(sparent[node - 1])
        if parent > nnodes:
            continue
        lo, hi = int(rptr[parent - 1]), int(rptr[parent])
        mp = {int(rlist[ii - 1]): ii - lo + 1 for ii in range(lo, hi)}
        for ii in range(int(rptr[node - 1]) + int(ncol[node - 1]), int(rptr[node])):
            out[ii - 1] = mp[int(rlist[ii - 1])]
    return out


def assign_nodes_to_levels(nnodes, sparent):
    level = [0] * (nnodes + 2)
    lvlcount = [0] * (nnodes + 2)
    num_levels = 1
    for node in range(nnodes, 0, -1):
        j = min(int(sparent[node - 1]), nnodes + 1)
        lvl = level[j] + 1
        level[node] = lvl
        lvlcount[lvl] += 1
        num_levels = max(num_levels, lvl)
    lvlptr = [0] * (num_levels + 2)
    lvlptr[1] = lvlptr[2] = 1
    for lvl in range(2, num_levels + 1):
        lvlptr[lvl + 1] = lvlptr[lvl] + lvlcount[num_levels - (lvl - 1) + 1]
    lvllist = [0] * nnodes
    for node in range(1, nnodes + 1):
        lvl = num_levels - level[node] + 1
        lvllist[lvlptr[lvl + 1] - 1] = node
        lvlptr[lvl + 1] += 1
    return num_levels, np.array(lvlptr[1:num_levels + 2]), np.array(lvllist)
