"""Profiling driver: analyse once, one warm-up factor, then one factor inside
cudaProfilerStart/Stop (use ncu --profile-from-start off)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import spral_b200 as sb
from spral_b200 import matrices as M

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 100
posdef = len(sys.argv) > 2 and sys.argv[2] == "posdef"
n, ptr, row, val = (M.laplacian_3d_7pt(grid) if posdef else M.stencil_3d_27pt(grid, shift=13.0))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from _order_cache import cached_metis_order
ak = sb.analyse(n, ptr, row, order=cached_metis_order(n, ptr, row))     # same ordering as order=None, METIS once per GPU call
dval = torch.from_numpy(val).cuda()
fk = sb.factor(ak, posdef, dval.data_ptr())
for ns in fk.numeric: ns.close()
torch.cuda.synchronize()
rt = torch.cuda.cudart()
from spral_b200 import _lib
if not os.environ.get("SPRAL_B200_NOPROFILE"):
    _lib.load().spral_ssids_b200_set_profile(1)      # NVTX range "upd_contrib" + event timing
solve_only = len(sys.argv) > 3 and sys.argv[3] == "solve"
if not solve_only:
    rt.cudaProfilerStart()
t = time.time()
fk = sb.factor(ak, posdef, dval.data_ptr())
torch.cuda.synchronize()
dt = time.time() - t
if not solve_only:
    rt.cudaProfilerStop()
tm = fk.numeric[0].timings()
print("factor", dt, "s", fk.inform["num_flops"] / dt / 1e9, "GF/s", "timings", tm[:8])
names = ["diag", "apply", "commit", "inner", "swap", "outer", "contrib(all)", "assemble", "init"]
print("class ms:", {k: round(float(v), 2) for k, v in zip(names, tm[8:17])}, "sum", round(float(tm[8:17].sum()), 1))
if len(sys.argv) > 3 and sys.argv[3] == "solve":
    nrhs = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    a = ak.analysis
    x = torch.ones(n * nrhs, dtype=torch.float64, device="cuda")
    for rep in range(2):
        if rep == 1:
            rt.cudaProfilerStart()
        torch.cuda.synchronize(); t = time.time()
        fk.numeric[0].solve_fwd(x.data_ptr(), nrhs, n)
        torch.cuda.synchronize(); t1 = time.time()
        fk.numeric[0].solve_diag_bwd(x.data_ptr(), nrhs, n)
        torch.cuda.synchronize(); t2 = time.time()
        if rep == 1:
            rt.cudaProfilerStop()
        print(f"solve nrhs={nrhs}: fwd {1e3*(t1-t):.2f} ms, diag+bwd {1e3*(t2-t1):.2f} ms")
