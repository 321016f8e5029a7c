"""Generates tests/golden/oracle_stats_cfg4.json: statistics of the reference CPU engine (oracle/_ref) on
BASELINE config 4 at benchmark size -- the structured KKT matrix matrices.kkt_grid(70) (n = 490 000, 30 %
zero-diagonal rows) with the matching-based scaling of spral_b200/scaling.py.  About a minute on 8 cores.
Run in the build container:  python tests/golden/make_golden_cfg4.py"""
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
from spral_b200 import matrices as M, ssids as host  # noqa: E402
from spral_b200.ssids import Analysis  # noqa: E402

if __name__ == "__main__":
    oracle_ref.ensure_env()
    g = 70
    n, ptr, row, val = M.kkt_grid(g)
    a = Analysis(n, ptr, row)
    s = host.compute_scaling(a, val, "hungarian")
    t = time.time()
    parts, r, sc = oracle_ref.ref_factor(a, False, val, scaling=s, nthreads=os.cpu_count())
    tf = time.time() - t
    A = M.to_scipy(n, ptr, row, val)
    b = np.asfortranarray(A @ np.ones((n, 1)))
    x = oracle_ref.ref_solve(a, parts, False, b, sc)
    out = dict(grid=g, n=int(n), constraints=int(n - g ** 3), nnodes=int(a.nnodes), scaling="hungarian",
               predicted_flops=int(a.num_flops), factor_seconds=tf, cores=os.cpu_count(),
               bwd=float(oracle_ref.backward_error(A, x, b)),
               **{k: int(r[k]) for k in ("flag", "num_delay", "num_neg", "num_two", "matrix_rank", "num_flops",
                                         "num_factor", "maxfront")})
    json.dump({"cfg4_kkt_grid70_hungarian": out}, open(os.path.join(HERE, "oracle_stats_cfg4.json"), "w"), indent=1,
              sort_keys=True)
    print(json.dumps(out, indent=1))
