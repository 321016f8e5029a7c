"""Developer check: factor a dense single-front matrix and verify P L D L^T P^T = A column by column."""
import os, sys, ctypes as C
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import spral_b200 as sb
from spral_b200 import matrices as M, _lib

def front(ns, node=0):
    lib = _lib.load()
    f = lib.spral_ssids_b200_debug_front
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    sz = np.zeros(5, np.int32)
    f(ns._h, node, sz.ctypes.data, None, None, None)
    m, n, ldl, nelim, ndin = map(int, sz)
    L = np.zeros(ldl * n); D = np.zeros(2 * n); perm = np.zeros(n, np.int32)
    f(ns._h, node, sz.ctypes.data, L.ctypes.data, D.ctypes.data, perm.ctypes.data)
    return m, n, nelim, L.reshape(n, ldl).T[:m, :], D, perm

def check(name, A):
    n, ptr, row, val = M._lower_csc_keep_zeros(sp.csc_matrix(A))
    order = np.arange(1, n + 1, dtype=np.int32)
    ak = sb.analyse(n, ptr, row, order=order)
    fk = sb.factor(ak, False, val)
    m, nn, nelim, L, D, perm = front(fk.numeric[0])
    Lm = np.tril(L[:, :nelim], -1) + np.eye(m, nelim)
    Dm = np.zeros((nelim, nelim))
    i = 0
    while i < nelim:
        if i + 1 < nelim and np.isinf(D[2 * i + 2]):
            Di = np.array([[D[2*i], D[2*i+1]], [D[2*i+1], D[2*i+3]]])
            Dm[i:i+2, i:i+2] = np.linalg.inv(Di); i += 2
        else:
            Dm[i, i] = 1.0 / D[2*i] if D[2*i] != 0 else 0.0; i += 1
    p = perm - 1
    Ap = A[np.ix_(p, p)]
    R = Lm @ Dm @ Lm.T - Ap
    colerr = np.abs(R).max(axis=0)
    bad = np.where(colerr > 1e-10)[0]
    print(f"[{name}] n={n} nelim={nelim} max|LDL^T-PAP^T|={np.abs(R).max():.2e} first bad col={bad[:1]} nbad={len(bad)}")
    if len(bad):
        c = bad[0]
        rows = np.where(np.abs(R[:, c]) > 1e-10)[0]
        print("   bad rows in first bad col:", rows[:10], "...", rows[-3:], "count", len(rows))
        # row-wise
        rowerr = np.abs(R).max(axis=1); badr = np.where(rowerr > 1e-10)[0]
        print("   bad rows overall:", badr[:10], "count", len(badr))
    sys.stdout.flush()

rng = np.random.default_rng(7)
def sym(n):
    A = rng.uniform(-1, 1, (n, n)); return (A + A.T) / 2
for n in (20, 33, 100, 257): sym(n)
check("dense-300", sym(300))
A = sym(130); A[:40, :40] *= 1e-6; check("smalllead-130", A)
A = sym(200); A[:60, :60] *= 1e-6; check("smalllead-200", A)
A = sym(300); A[:100, :100] *= 1e-6; check("smalllead-300", A)
