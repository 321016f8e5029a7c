"""TEST INFRASTRUCTURE (CPU only): the one-process-per-GPU driver (spral_b200/dist.py) on its REAL GPU code path --
GpuEngine, contribution blocks handed over by "CUDA IPC", the distributed top front (SPRAL_B200_SPLIT=1: owner rank and
helper rank talking through POSIX shared memory) -- as two real processes over gloo, on the SIMT emulator of tests/emu
with shared-memory-backed "device" memory (SPRAL_B200_EMU_SHM=1), so that an IPC handle can be opened by the other
process.  The only thing injected is where the right-hand sides live (torch has no CUDA device here).
usage: dist_split_check.py [grid=28] [stencil|lap] [ranks=2]   prints one JSON line; exit code 0 when everything agrees."""
import json
import multiprocessing as mp
import os
import socket
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "build", "emu_split", "libspral_ssids_b200_emu_split.so")


def _worker(rank, world, port, grid, kind, logdir, q):
    try:
        log = open(os.path.join(logdir, f"rank{rank}.log"), "w")
        os.dup2(log.fileno(), 2)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2", OMP_CANCELLATION="TRUE",
                          SPRAL_B200_EMU_SHM="1", SPRAL_B200_SPLIT="1", SPRAL_B200_TRACE="1",
                          SPRAL_B200_SPLIT_TIMEOUT="120")
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import numpy as np
        import torch
        import torch.distributed as dist
        from spral_b200 import _lib
        _lib.LIB_PATH = LIB
        import spral_b200 as sb
        from spral_b200 import matrices as M, dist as sdist
        import oracle_ref

        class EmuGpuEngine(sdist.GpuEngine):
            def device(self, local_rank):
                return torch.device("cpu")

        dist.init_process_group("gloo", rank=rank, world_size=world)
        posdef = kind == "lap"
        n, ptr, row, val = M.laplacian_3d_7pt(grid) if posdef else M.stencil_3d_27pt(grid, shift=13.0)
        ctx = sdist.DistContext(world, rank, 0, engine=EmuGpuEngine())
        ak = sdist.analyse(ctx, n, ptr, row)
        A = M.to_scipy(n, ptr, row, val)
        rng = np.random.default_rng(3)
        B = np.asfortranarray(A @ rng.uniform(-1, 1, (n, 2)))
        out = {}
        for rep in range(2):                         # twice: the shared-memory names follow the epoch
            fk = sdist.factor(ctx, ak, posdef, val)
            inform = sdist.reduce_inform(ctx, fk.inform)
            X = sdist.solve(ctx, fk, B)
            out = dict(inform={k: int(inform[k]) for k in ("flag", "num_neg", "matrix_rank", "num_delay", "num_factor", "num_flops")},
                       bwd=float(oracle_ref.backward_error(A, X, B)), owners=[int(r) for r in ak.rank_of], nparts=int(ak.analysis.nparts))
            sdist.free(fk)
        if rank == 0:                                # the same library, one process, no split
            ak1 = sb.analyse(n, ptr, row)
            fk1 = sb.factor(ak1, posdef, val)
            X1 = sb.solve(fk1, B)
            g = fk1.inform
            out["single"] = {k: int(g[k]) for k in ("flag", "num_neg", "matrix_rank", "num_delay", "num_factor", "num_flops")}
            out["maxdiff"] = float(np.abs(X1 - X).max() / max(1.0, np.abs(X1).max()))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception:
        import traceback
        q.put((rank, dict(error=traceback.format_exc())))


def main():
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    kind = sys.argv[2] if len(sys.argv) > 2 else "stencil"       # stencil: 27-point indefinite; lap: 7-point Laplacian, Cholesky
    world = int(sys.argv[3]) if len(sys.argv) > 3 else 2             # ranks: one owner of the top fronts, world - 1 helpers
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    with tempfile.TemporaryDirectory() as logdir:
        procs = [ctxm.Process(target=_worker, args=(r, world, port, grid, kind, logdir, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = dict(q.get(timeout=1700) for _ in procs)
        for p in procs:
            p.join(timeout=60)
        logs = "".join(open(os.path.join(logdir, f"rank{r}.log")).read() for r in range(world))
    split_lines = [l for l in logs.splitlines() if l.startswith("[split]") or "split helper returned" in l]
    ok = all("error" not in res[r] for r in res)
    if ok:
        r0 = res[0]
        ok = (r0["inform"] == r0["single"] and r0["bwd"] < 5e-11 and r0["maxdiff"] < 1e-9 and all(res[r]["inform"] == r0["inform"] for r in res)
              and any("panels pushed" in l and "front closed" in l and not l.startswith("[split] front closed (0 ") for l in split_lines)
              and sum("split helper returned 0" in l for l in split_lines) >= 2 * (world - 1))
    print(json.dumps(dict(ok=ok, res=res, split=split_lines[-8:], log=None if ok else logs[-3000:])))
    for f in os.listdir("/dev/shm"):
        if f.startswith("spral_emu_") and any(f.startswith(f"spral_emu_{p.pid}_") for p in procs):
            os.unlink(os.path.join("/dev/shm", f))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
