"""Factorises a few fixed matrices on cuda:0 and dumps what identifies the result
bit for bit (pivot order, D^-1, inform, one solution) to an .npz.  Used by
tests/test_gpu_paths.py to compare an alternative code path (selected
through the SPRAL_B200_* environment of THIS process) with the default engine."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("SPRAL_B200_EMU_LIB"):      # developer / test hook: the CPU-emulated build of tests/emu (see tests/conftest.py)
    from spral_b200 import _lib as _emu_lib
    _emu_lib.LIB_PATH = os.environ["SPRAL_B200_EMU_LIB"]
import spral_b200 as sb                      # noqa: E402
from spral_b200 import matrices as M         # noqa: E402

def _dense_sym(n, posdef, seed):
    """One dense front: the whole matrix is the root supernode (exercises the panel kernels alone)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (n, n))
    A = (A + A.T) / 2
    if posdef:
        A = A @ A.T / n + np.eye(n)
    return M._lower_csc_keep_zeros(sp.csc_matrix(A))


CASES = {
    "dense_600_indef": (lambda: _dense_sym(600, False, 1), False),
    "dense_391_indef": (lambda: _dense_sym(391, False, 2), False),
    "dense_500_posdef": (lambda: _dense_sym(500, True, 3), True),
    "stencil27_20_indef": (lambda: M.stencil_3d_27pt(20, shift=13.0), False),
    "stencil27_36_indef": (lambda: M.stencil_3d_27pt(36, shift=13.0), False),
    "lap3d_24_posdef": (lambda: M.laplacian_3d_7pt(24), True),
    "kkt_3000": (lambda: M.kkt_saddle(3000, 0.3, seed=3), False),
}


def main(out):
    res = {}
    only = os.environ.get("SPRAL_B200_DUMP_CASES")          # comma-separated subset (the emulated build is slow)
    for name, (gen, posdef) in CASES.items():
        if only and name not in only.split(","):
            continue
        n, ptr, row, val = gen()
        ak = sb.analyse(n, ptr, row)
        fk = sb.factor(ak, posdef, val)
        piv, d = fk.numeric[0].enquire()
        A = M.to_scipy(n, ptr, row, val)
        x = sb.solve(fk, A @ np.ones(n))
        rng = np.random.default_rng(11)
        X5 = sb.solve(fk, np.asfortranarray(A @ rng.uniform(-1, 1, (n, 5))))      # chunks of 4 + 1 right-hand sides
        res[name + "/x5"] = X5
        res[name + "/x20"] = sb.solve(fk, np.asfortranarray(A @ rng.uniform(-1, 1, (n, 20))))      # 16 (tensor cores) + 4
        g = fk.inform
        res[name + "/d"] = d
        if piv is not None:
            res[name + "/piv"] = piv
        res[name + "/x"] = x
        res[name + "/inform"] = np.array([g[k] for k in ("flag", "num_delay", "num_neg", "num_two", "matrix_rank",
                                                          "num_factor", "num_flops")], dtype=np.int64)
    np.savez(out, **res)


if __name__ == "__main__":
    main(sys.argv[1])
