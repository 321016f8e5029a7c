"""Generates tests/golden/oracle_stats.json: statistics of the reference CPU
engine (oracle/_ref, compiled from /root/reference) on small synthetic
matrices.  Run in the build container:  python tests/golden/make_golden.py"""
import json
import os
import sys

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
    "example_5x5": (M.example_5x5, False),
    "lap2d_100": (lambda: M.laplacian_2d_5pt(100), True),
    "lap3d_14": (lambda: M.laplacian_3d_7pt(14), True),
    "st27_12_s13": (lambda: M.stencil_3d_27pt(12, shift=13.0), False),
    "st27_16_s5": (lambda: M.stencil_3d_27pt(16, shift=5.0), False),
    "kkt_1500": (lambda: M.kkt_saddle(1500), False),
}


def run_case(name):
    gen, posdef = CASES[name]
    n, ptr, row, val = gen()
    a = Analysis(n, ptr, row)
    parts, inform, sc = oracle_ref.ref_factor(a, posdef, val, nthreads=1)
    A = M.to_scipy(n, ptr, row, val)
    b = A @ np.ones(n)
    x = oracle_ref.ref_solve(a, parts, posdef, b)
    out = dict(n=int(n), nnodes=int(a.nnodes), num_factor=int(inform["num_factor"]),
               num_flops=int(inform["num_flops"]), maxfront=int(inform["maxfront"]),
               num_neg=int(inform["num_neg"]), num_two=int(inform["num_two"]),
               num_delay=int(inform["num_delay"]), matrix_rank=int(inform["matrix_rank"]),
               bwd=float(oracle_ref.backward_error(A, x, b)))
    for p in parts:
        p.close()
    a.close()
    return out


if __name__ == "__main__":
    oracle_ref.ensure_env()
    gold = {k: run_case(k) for k in CASES}
    json.dump(gold, open(os.path.join(HERE, "oracle_stats.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(gold, indent=1))
