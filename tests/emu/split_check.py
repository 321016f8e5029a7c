"""TEST INFRASTRUCTURE (CPU only): the distributed top front END TO END on the SIMT emulator -- the library built with
-DSPRAL_B200_SPLIT (tests/emu/build_emu.py split), the owner's factorisation in this thread, the helper's service loop
(spral_ssids_b200_split_helper_serve) in another, one dense front.  The factors and the solution must equal the ones of
the same library without a helper, bit for bit.  usage: split_check.py n indef|posdef|saddle [helpers=1]"""
import sys, os, time, threading, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ.setdefault("OMP_CANCELLATION", "TRUE")
import numpy as np, scipy.sparse as sp
from spral_b200 import _lib
_lib.LIB_PATH = os.environ['SPRAL_B200_EMU_LIB']
import spral_b200 as sb
from spral_b200 import matrices as M
import oracle_ref
lib = _lib.load()
lib.spral_ssids_b200_split_enable.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
lib.spral_ssids_b200_split_helper_serve.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int]
lib.spral_ssids_b200_split_helper_serve.restype = C.c_int
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1300
kind = sys.argv[2] if len(sys.argv) > 2 else "indef"
NH = int(sys.argv[3]) if len(sys.argv) > 3 else 1          # helper threads
rng = np.random.default_rng(n)
A = rng.uniform(-1, 1, (n, n)); A = (A + A.T) / 2
posdef = kind == "posdef"
if posdef: A = A @ A.T / n + np.eye(n)
if kind == "saddle": A[:n // 3, :n // 3] = 0.0
n_, ptr, row, val = M._lower_csc_keep_zeros(sp.csc_matrix(A))
order = np.arange(1, n + 1, dtype=np.int32)
As = M.to_scipy(n_, ptr, row, val)
B = np.asfortranarray(As @ rng.uniform(-1, 1, (n, 2)))
res = {}
for split in (False, True):
    ak = sb.analyse(n_, ptr, row, order=order)
    name = f"/spral_b200_emu_split_{os.getpid()}".encode()
    rc = [None]
    if split:
        lib.spral_ssids_b200_split_enable(ak.subtrees[-1]._h, name, NH)
        rcs = [None] * NH
        ths = [threading.Thread(target=lambda h=h: rcs.__setitem__(h, lib.spral_ssids_b200_split_helper_serve(name, 0, 30.0, h))) for h in range(NH)]
        for th in ths: th.start()
    t = time.time()
    fk = sb.factor(ak, posdef, val)
    dt = time.time() - t
    if split:
        for th in ths: th.join()
        rc[0] = max(rcs)
    X = sb.solve(fk, B)
    piv, d = fk.numeric[0].enquire()
    res[split] = (fk.inform, X, d)
    print("split" if split else "whole", f"{dt:.1f}s helper rc {rc[0]}", {k: fk.inform[k] for k in ("flag", "num_neg", "matrix_rank", "num_delay", "num_two", "not_first_pass")},
          "bwd %.2e" % oracle_ref.backward_error(As, X, B))
same = np.array_equal(res[False][1], res[True][1]) and np.array_equal(res[False][2], res[True][2], equal_nan=True)
print("BITWISE IDENTICAL" if same else "DIFFERENT", "max |dx|", float(np.abs(res[False][1] - res[True][1]).max()))
sys.exit(0 if same and rc[0] == 0 else 1)
