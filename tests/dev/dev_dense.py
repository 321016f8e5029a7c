"""Developer check: dense single-front matrices that force pivot failures."""
import os, sys, traceback
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_ref
oracle_ref.ensure_env()
import spral_b200 as sb
from spral_b200 import matrices as M


def lower(A):
    return M._lower_csc_keep_zeros(sp.csc_matrix(A))


def run(name, A, posdef=False, nemin=32):
    try:
        n, ptr, row, val = lower(A)
        order = np.arange(1, n + 1, dtype=np.int32)
        ak = sb.analyse(n, ptr, row, order=order, nemin=nemin)
        a = ak.analysis
        As = M.to_scipy(n, ptr, row, val)
        b = As @ np.ones(n)
        fk = sb.factor(ak, posdef, val)
        x = sb.solve(fk, b)
        inf = fk.inform
        be = oracle_ref.backward_error(As, x, b)
        parts, rinf, sc = oracle_ref.ref_factor(a, posdef, val)
        xr = oracle_ref.ref_solve(a, parts, posdef, b)
        keys = ('flag', 'num_delay', 'num_neg', 'num_two', 'matrix_rank', 'not_first_pass', 'not_second_pass')
        print(f"[{name}] n={n} nnodes={a.nnodes} maxfront={a.maxfront} bwd={be:.2e} ref_bwd={oracle_ref.backward_error(As, xr, b):.2e}")
        print("   gpu:", {k: inf[k] for k in keys})
        print("   ref:", {k: rinf[k] for k in keys})
        for p_ in parts: p_.close()
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()


rng = np.random.default_rng(7)
def sym(n):
    A = rng.uniform(-1, 1, (n, n)); return (A + A.T) / 2

for n in (20, 33, 100, 257, 300, 600):
    run(f"dense-indef-{n}", sym(n))
# tiny leading diagonal: forces failures / 2x2 across blocks
for n in (64, 100, 300, 600):
    A = sym(n); k = n // 3
    A[:k, :k] *= 1e-6
    run(f"dense-smalllead-{n}", A)
# saddle point, dense: zero (1,1) block
for n in (64, 100, 300, 600):
    A = sym(n); k = n // 3
    A[:k, :k] = 0.0
    run(f"dense-saddle-{n}", A)
# two-front problem: leading block coupled to a trailing dense block through rows below
for n in (200, 700):
    A = sym(n); k = n // 2
    A[:k, :k] = np.diag(np.diag(A[:k, :k])) * 1e-8   # nearly zero diagonal leading part, no coupling inside
    run(f"arrow-{n}", A, nemin=1)
