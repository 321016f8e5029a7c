#!/usr/bin/env python
"""Command-line driver in the image of the reference's driver/spral_ssids.F90: reads a
Rutherford-Boeing matrix (default matrix.rb), rhs = A * 1, analyse / factor / solve on
the B200 engine, prints the timing lines, the forward error, the scaled backward error
(driver/spral_ssids.F90:419-480) and the statistics block (:195-209).

  python tools/spral_ssids.py [file.rb] [--posdef] [--nrhs K] [--nemin N] [--u U]
        [--scale=none|mc64|auction|mc77] [--ordering=mc64-metis] [--gen NAME]

--gen NAME uses a synthetic matrix instead of a file: lap2d:<g>, lap3d:<g>, st27:<g>[:shift],
kkt_grid:<g>.  A CUDA device is required (there is no CPU fallback)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def generate(spec):
    from spral_b200 import matrices as M
    name, *a = spec.split(":")
    if name == "lap2d":
        return M.laplacian_2d_5pt(int(a[0]))
    if name == "lap3d":
        return M.laplacian_3d_7pt(int(a[0]))
    if name == "st27":
        return M.stencil_3d_27pt(int(a[0]), shift=float(a[1]) if len(a) > 1 else 13.0)
    if name == "kkt_grid":
        return M.kkt_grid(int(a[0]))
    raise SystemExit(f"unknown generator {spec}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("filename", nargs="?", default="matrix.rb")
    ap.add_argument("--gen")
    ap.add_argument("--posdef", action="store_true")
    ap.add_argument("--nrhs", type=int, default=1)
    ap.add_argument("--nemin", type=int, default=32)
    ap.add_argument("--u", type=float, default=0.01)
    ap.add_argument("--scale", default="none", choices=["none", "mc64", "auction", "mc77"])
    ap.add_argument("--ordering", default="metis", choices=["metis", "mc64-metis"])
    args = ap.parse_args()

    import spral_b200 as sb
    from spral_b200 import matrices as M, rb, _lib

    if args.gen:
        n, ptr, row, val = generate(args.gen)
        print(f"Generated '{args.gen}'")
    else:
        print(f"Reading '{args.filename}'...", end="")
        n, ptr, row, val, info = rb.rb_read(args.filename)
        print("ok")
    A = M.to_scipy(n, ptr, row, val)
    rhs = np.asfortranarray(np.repeat((A @ np.ones(n))[:, None], args.nrhs, axis=1))

    opt = _lib.Options.default()
    opt.u = args.u
    matching = args.ordering == "mc64-metis"
    print("The computed solution...")
    t = time.perf_counter()
    ak = sb.analyse(n, ptr, row, nemin=args.nemin, options=opt, val=val if matching else None,
                    ordering="matching" if matching else None)
    a = ak.analysis
    print("ok")
    t_anal = time.perf_counter() - t
    print(f" Analyse took  {t_anal:.4f}")
    print(f"Predict nfact = {a.num_factor:10.2e}")
    print(f"Predict nflop = {a.num_flops:10.2e}")
    print(f"nparts{a.nparts:10d}")
    scaling = {"none": None, "mc64": "hungarian", "auction": "auction", "mc77": "equilib"}[args.scale]
    if matching:
        scaling = "matching"
    print("Factorize...")
    t = time.perf_counter()
    fk = sb.factor(ak, args.posdef, val, options=opt, scaling=scaling)
    t_fact = time.perf_counter() - t
    g = fk.inform
    if g["flag"] < 0:
        print(" oops on factorize ", g["flag"])
        return 1
    print("ok")
    print(f" Factor took  {t_fact:.4f}   ({g['num_flops'] / t_fact / 1e9:.1f} GFLOP/s, device "
          f"{sum(float(ns.timings()[1]) for ns in fk.numeric):.1f} ms)")
    print("Solve...")
    t = time.perf_counter()
    soln = sb.solve(fk, rhs)
    t_solve = time.perf_counter() - t
    print("ok")
    print(f" Solve took  {t_solve:.4f}")
    soln = soln.reshape(n, -1)
    print(" number bad cmp = ", int((np.abs(soln[:, 0] - 1.0) >= 1e-6).sum()))
    print(" fwd error || ||_inf = ", float(np.abs(soln[:, 0] - 1.0).max()))
    r = A @ soln - rhs
    anorm = abs(A).sum(axis=1).max()
    res = (np.abs(r).max(axis=0) / (anorm * np.abs(soln).max(axis=0) + np.abs(rhs).max(axis=0))).max()
    print(" bwd error scaled = ", float(res))
    print(f"{'cmp:':>6}{'SMFCT':>10}")
    print(f"{'anal:':>6}{t_anal:10.2f}")
    print(f"{'fact:':>6}{t_fact:10.2f}")
    print(f"{'afact:':>6}{g['num_factor']:10.2e}")
    print(f"{'aflop:':>6}{g['num_flops']:10.2e}")
    print(f"{'nfact:':>6}{a.num_factor:10.2e}")
    print(f"{'nflop:':>6}{a.num_flops:10.2e}")
    print(f"{'delay:':>6}{g['num_delay']:10d}")
    print(f"{'inertia:':>6}{g['num_neg']:10d}{n - g['matrix_rank']:10d}{g['matrix_rank'] - g['num_neg']:10d}")
    print(f"{'2x2piv:':>6}{g['num_two']:10d}")
    print(f"{'maxfront:':>6}{g['maxfront']:10d}")
    print(f"{'maxsupernode:':>6}{g['maxsupernode']:10d}")
    print(f"{'not_first_pass:':>6}{g['not_first_pass']:10d}")
    print(f"{'not_second_pass:':>6}{g['not_second_pass']:10d}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
