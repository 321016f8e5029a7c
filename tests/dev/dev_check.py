"""Developer smoke run on a GPU box: a series of small factor/solve cases
against the reference CPU engine (oracle/_ref).  Not part of the test-suite."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_ref
oracle_ref.ensure_env()
import spral_b200 as sb
from spral_b200 import matrices as M
from spral_b200.ssids import Analysis


def run(name, gen, posdef, nrhs=1, ngpu=1, check_ref=True, **akw):
    try:
        n, ptr, row, val = gen()
        t = time.time()
        ak = sb.analyse(n, ptr, row, ngpu=ngpu, **akw)
        a = ak.analysis
        ta = time.time() - t
        A = M.to_scipy(n, ptr, row, val)
        rng = np.random.default_rng(1)
        X = np.asfortranarray(rng.uniform(-1, 1, size=(n, nrhs)))
        X[:, 0] = 1.0
        B = np.asfortranarray(A @ X)
        t = time.time(); fk = sb.factor(ak, posdef, val); tf = time.time() - t
        t = time.time(); fk2 = sb.factor(ak, posdef, val); tf2 = time.time() - t
        inf = fk2.inform
        t = time.time(); Xs = sb.solve(fk2, B); ts = time.time() - t
        be = oracle_ref.backward_error(A, Xs, B)
        print(f"[{name}] n={n} nnodes={a.nnodes} nparts={a.nparts} lev? t_an={ta:.2f} t_f={tf:.3f}/{tf2:.3f} t_s={ts:.3f} "
              f"GF/s={inf['num_flops']/tf2/1e9:.1f} bwd={be:.2e}")
        print("   gpu:", {k: inf[k] for k in ('flag','num_delay','num_neg','num_two','matrix_rank','num_factor','num_flops','maxfront','not_first_pass','not_second_pass')})
        print("   timings:", [round(float(x),2) for x in fk2.numeric[-1].timings()[:6]])
        if check_ref:
            parts, rinf, sc = oracle_ref.ref_factor(a, posdef, val)
            xr = oracle_ref.ref_solve(a, parts, posdef, B)
            print("   ref:", {k: rinf[k] for k in ('flag','num_delay','num_neg','num_two','matrix_rank','num_factor','num_flops','maxfront','not_first_pass','not_second_pass')},
                  f"t_f={rinf['factor_time']:.3f} bwd={oracle_ref.backward_error(A, xr, B):.2e}")
            for p_ in parts: p_.close()
        sys.stdout.flush()
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()


if __name__ == "__main__":
    which = sys.argv[1:] or ["small"]
    if "small" in which:
        run("5x5", M.example_5x5, False)
        run("lap2d-10", lambda: M.laplacian_2d_5pt(10), True)
        run("lap2d-100", lambda: M.laplacian_2d_5pt(100), True, nrhs=3)
        run("lap2d-100-indef", lambda: M.laplacian_2d_5pt(100), False)
        run("27pt-12-indef", lambda: M.stencil_3d_27pt(12, shift=13.0), False, nrhs=2)
        run("27pt-20-indef", lambda: M.stencil_3d_27pt(20, shift=13.0), False, nrhs=5)
        run("lap3d-20", lambda: M.laplacian_3d_7pt(20), True, nrhs=9)
        run("kkt-2000", lambda: M.kkt_saddle(2000), False)
    if "mid" in which:
        run("lap3d-40", lambda: M.laplacian_3d_7pt(40), True)
        run("27pt-40-indef", lambda: M.stencil_3d_27pt(40, shift=13.0), False)
    if "cfg2" in which:
        run("lap3d-60", lambda: M.laplacian_3d_7pt(60), True)
    if "cfg3" in which:
        run("27pt-80", lambda: M.stencil_3d_27pt(80, shift=13.0), False, check_ref=False)
    if "cfg5" in which:
        run("27pt-100", lambda: M.stencil_3d_27pt(100, shift=13.0), False, check_ref=False)
    if "multi" in which:
        run("27pt-20-2parts", lambda: M.stencil_3d_27pt(20, shift=13.0), False, ngpu=2, devices=[0, 0])
        run("lap3d-20-4parts", lambda: M.laplacian_3d_7pt(20), True, ngpu=4, devices=[0, 0, 0, 0])
