"""Generates tests/golden/oracle_stats_full.json: statistics of the reference CPU engine (oracle/_ref,
compiled from /root/reference) on the BASELINE configs at FULL size --
  cfg2  3-D 7-point Laplacian 60^3, positive definite
  cfg3  3-D 27-point 80^3, shifted (sigma = 13) indefinite, u = 0.01
  cfg5  3-D 27-point 100^3, shifted (sigma = 13) indefinite, u = 0.01   (the headline workload)
same ordering (METIS through the restated analyse) and options as the GPU tests use.  A few minutes on 8
cores.  Run in the build container:  python tests/golden/make_golden_full.py [cfg2 cfg3 cfg5]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import oracle_ref  # noqa: E402
from spral_b200 import matrices as M  # noqa: E402
from spral_b200.ssids import Analysis  # noqa: E402

CASES = {
    "cfg2": (lambda: M.laplacian_3d_7pt(60), True),
    "cfg3": (lambda: M.stencil_3d_27pt(80, shift=13.0), False),
    "cfg5": (lambda: M.stencil_3d_27pt(100, shift=13.0), False),
}


def run_case(name):
    gen, posdef = CASES[name]
    n, ptr, row, val = gen()
    a = Analysis(n, ptr, row)
    cores = os.cpu_count()
    t = time.time()
    parts, r, _ = oracle_ref.ref_factor(a, posdef, val, nthreads=cores)
    tf = time.time() - t
    A = M.to_scipy(n, ptr, row, val)
    rng = np.random.default_rng(0)
    X = np.asfortranarray(rng.uniform(-1, 1, (n, 2)))
    X[:, 0] = 1.0
    B = np.asfortranarray(A @ X)
    t = time.time()
    Xs = oracle_ref.ref_solve(a, parts, posdef, B)
    ts = time.time() - t
    bwd = float(oracle_ref.backward_error(A, Xs, B))
    Xs2 = Xs + oracle_ref.ref_solve(a, parts, posdef, np.asfortranarray(B - A @ Xs))
    bwd2 = float(oracle_ref.backward_error(A, Xs2, B))
    out = dict(n=int(n), nnodes=int(a.nnodes), nparts=int(a.nparts), flag=int(r["flag"]),
               predicted_num_factor=int(a.num_factor), predicted_num_flops=int(a.num_flops),
               num_factor=int(r["num_factor"]), num_flops=int(r["num_flops"]), maxfront=int(r["maxfront"]),
               num_neg=int(r["num_neg"]), num_two=int(r["num_two"]), num_delay=int(r["num_delay"]),
               matrix_rank=int(r["matrix_rank"]), bwd=bwd, bwd_after_1_refinement=bwd2,
               factor_seconds=tf, solve_2rhs_seconds=ts, cores=cores)
    for p in parts:
        p.close()
    a.close()
    return out


if __name__ == "__main__":
    oracle_ref.ensure_env()
    path = os.path.join(HERE, "oracle_stats_full.json")
    gold = json.load(open(path)) if os.path.exists(path) else {}
    for k in (sys.argv[1:] or list(CASES)):
        gold[k] = run_case(k)
        print(k, json.dumps(gold[k]), flush=True)
        json.dump(gold, open(path, "w"), indent=1, sort_keys=True)
