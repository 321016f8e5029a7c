"""A/B of the opt-in kernel variants on one GPU: for every variant a fresh process
(the library reads its SPRAL_B200_* switches once) factorises the stencil problem a few
times and reports the best device time, the class times of a profiled run, inertia /
delays and the solve times.  One table at the end.

  python tools/ab_variants.py [grid=100] [reps=3] [variant ...]
  variants: base diag_v2 panel_v2 panel_v2+diag... any '+'-joined combination of
            diag_v2 panel_v2 bulk_prio bulk84 ctile8 ctile12 solve_wide
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = {
    "base": {},
    "bulk100": {"SPRAL_B200_BULK_CTAS": "100"},
    "bulk132": {"SPRAL_B200_BULK_CTAS": "132"},
    "v2f64": {"SPRAL_B200_PANEL_V2_FRONTS": "64"},
    "v2f8": {"SPRAL_B200_PANEL_V2_FRONTS": "8"},
    "panel_v2": {"SPRAL_B200_PANEL_V2": "1"},
    "bulk_prio": {"SPRAL_B200_BULK_PRIO": "1"},
    "ctile12": {"SPRAL_B200_CTILE_BLOCK": "12"},
    "ctile8": {"SPRAL_B200_CTILE_BLOCK": "8"},       # 16 operand panels of 5.4 MB (K = 5243) stay inside the 126 MB L2
    "bulk84": {"SPRAL_B200_BULK_CTAS": "84"},        # more SMs left for the panel kernels (PANEL_V2 tiles: 1 CTA / SM)
    "solve_wide": {"SPRAL_B200_SOLVE_WIDE": "1"},
    "solve_nolookahead": {"SPRAL_B200_SOLVE_LOOKAHEAD": "0"},
    "solve_nolinv": {"SPRAL_B200_SOLVE_LINV": "0"},
}
DEFAULT = ["base", "diag1", "panel_v2", "panel_v2+bulk_prio", "panel_v2+bulk100", "panel_v2+bulk132", "panel_v2+v2f8",
           "panel_v2+bulk_prio+solve_wide"]


def child(grid, reps):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import spral_b200 as sb
    from spral_b200 import matrices as M, _lib
    n, ptr, row, val = M.stencil_3d_27pt(grid, shift=13.0)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from _order_cache import cached_metis_order
    ak = sb.analyse(n, ptr, row, order=cached_metis_order(n, ptr, row))     # same ordering, METIS once per GPU call
    dval = torch.from_numpy(val).cuda()
    best, fk = 1e30, None
    for _ in range(reps + 1):
        if fk is not None:
            for ns in fk.numeric:
                ns.close()
        fk = sb.factor(ak, False, dval.data_ptr())
        best = min(best, float(fk.numeric[0].timings()[1]))
    g = fk.inform
    A = M.to_scipy(n, ptr, row, val)
    b = A @ np.ones(n)
    X = torch.from_numpy(np.ascontiguousarray(b)).cuda()
    ns = fk.numeric[0]
    # the solve entry points take pivot-order vectors; a backward-error check goes through sb.solve
    x = sb.solve(fk, b)
    r = A @ x - b                                     # scaled residual of driver/spral_ssids.F90:419-480
    be = float(np.abs(r).max() / (abs(A).sum(axis=1).max() * np.abs(x).max() + np.abs(b).max()))
    ts = {}
    for nrhs in (1, 32):
        xx = torch.ones(n * nrhs, dtype=torch.float64, device="cuda")
        for rep in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            ns.solve_fwd(xx.data_ptr(), nrhs, n)
            ns.solve_diag_bwd(xx.data_ptr(), nrhs, n)
            torch.cuda.synchronize(); ts[nrhs] = 1e3 * (time.perf_counter() - t)
    lib = _lib.load()
    lib.spral_ssids_b200_set_profile(1)
    for q in fk.numeric:
        q.close()
    fk = sb.factor(ak, False, dval.data_ptr())
    lib.spral_ssids_b200_set_profile(0)
    tm = fk.numeric[0].timings()
    names = ["diag", "apply", "commit", "inner", "swap", "outer", "contrib", "assemble", "init"]
    out = {"ms": best, "tflops": g["num_flops"] / best / 1e9, "launches": int(tm[6]),
           "num_neg": int(g["num_neg"]), "num_delay": int(g["num_delay"]), "num_two": int(g["num_two"]),
           "rank": int(g["matrix_rank"]), "bwd": be, "solve1_ms": ts[1], "solve32_ms": ts[32],
           "class_ms": {k: round(float(v), 2) for k, v in zip(names, tm[8:17])},
           "contrib_tflops": float(tm[3]) / (float(tm[2]) * 1e-3) / 1e12 if tm[2] > 0 else 0.0}
    print("AB_RESULT " + json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(int(sys.argv[2]), int(sys.argv[3]))
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    variants = sys.argv[3:] or DEFAULT
    rows = []
    for v in variants:
        env = dict(os.environ)
        for part in v.split("+"):
            env.update(ENV[part])
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(grid), str(reps)], env=env,
                               capture_output=True, text=True, timeout=1500)
            line = [l for l in r.stdout.splitlines() if l.startswith("AB_RESULT ")]
            res = json.loads(line[-1][10:]) if line else {"error": (r.stderr or r.stdout)[-400:]}
        except subprocess.TimeoutExpired:
            res = {"error": "timeout"}
        rows.append((v, res))
        print(v, json.dumps(res), flush=True)
    print("\n| variant | factor ms | TFLOP/s | launches | num_neg | delays | bwd err | solve 1 / 32 RHS ms | chain ms (diag+apply+commit+inner+swap) | outer | contrib TF/s |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for v, r in rows:
        if "error" in r:
            print(f"| {v} | ERROR: {r['error'][:120]!r} |")
            continue
        c = r["class_ms"]
        chain = c["diag"] + c["apply"] + c["commit"] + c["inner"] + c["swap"]
        print(f"| {v} | {r['ms']:.1f} | {r['tflops']:.2f} | {r['launches']} | {r['num_neg']} | {r['num_delay']} | {r['bwd']:.1e} | "
              f"{r['solve1_ms']:.1f} / {r['solve32_ms']:.1f} | {chain:.1f} | {c['outer']:.1f} | {r['contrib_tflops']:.1f} |")


if __name__ == "__main__":
    main()
