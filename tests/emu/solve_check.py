"""TEST INFRASTRUCTURE (CPU): the solve paths that only many right-hand sides or large fronts reach, on the emulated
library (tests/test_emulated_engine.py sets SPRAL_B200_EMU_LIB): tensor-core G kernels (16+ right-hand sides), the
look-ahead inside a sweep (fronts of 2+ blocks of 256 columns), the inverse diagonal blocks (fronts of 256+ columns,
L D L^T and Cholesky), the half-height T kernels (levels of small fronts), chunks of 64 + 16 + 4 right-hand sides.
Usage: solve_check.py  -> prints one line per case, exits 1 on a wrong solution."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: F401  (points the loader at the emulated library)
import numpy as np
import scipy.sparse as sp

import spral_b200 as sb
from spral_b200 import matrices as M


def dense(n, posdef, seed):
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (n, n))
    A = A @ A.T + n * np.eye(n) if posdef else A + A.T
    return M._lower_csc_keep_zeros(sp.csc_matrix(A))


CASES = [
    ("dense 300 indef, 16 rhs", lambda: dense(300, False, 5), False, 16, 1e-9),
    ("dense 300 posdef, 20 rhs", lambda: dense(300, True, 6), True, 20, 1e-12),
    ("dense 520 indef, 1 rhs", lambda: dense(520, False, 7), False, 1, 1e-9),
    ("stencil27 10^3, 84 rhs", lambda: M.stencil_3d_27pt(10, shift=13.0), False, 84, 1e-10),
    ("laplacian 12^3 posdef, 33 rhs", lambda: M.laplacian_3d_7pt(12), True, 33, 1e-12),
]


def main():
    bad = 0
    for name, gen, posdef, nrhs, tol in CASES:
        n, ptr, row, val = gen()
        fk = sb.factor(sb.analyse(n, ptr, row), posdef, val)
        A = M.to_scipy(n, ptr, row, val)
        rng = np.random.default_rng(1)
        X = np.asfortranarray(rng.uniform(-1, 1, (n, nrhs)))
        B = np.asfortranarray(A @ X)
        Xs = sb.solve(fk, B if nrhs > 1 else B[:, 0]).reshape(n, -1)
        err = float(np.abs(Xs - X).max())
        ok = fk.inform["flag"] >= 0 and err <= tol
        print(f"{name}: max error {err:.2e} {'OK' if ok else 'WRONG'}")
        bad += not ok
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
