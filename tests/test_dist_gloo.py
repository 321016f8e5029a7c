"""World-size-2 (and 3) gloo tests on CPU of the one-process-per-GPU driver
(spral_b200/dist.py): part ownership, cross-rank contribution hand-over, the
solve exchanges and the final all-reduce.  The reference CPU engine stands in
for the GPU engine through an injected adapter (tests/oracle_engine.py), so what is
tested here is the host logic the GPU path shares."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle_ref
from conftest import ROOT

pytestmark = pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch.distributed as dist
        import oracle_ref as ref
        ref.ensure_env()
        from spral_b200 import matrices as M, dist as sdist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        gen, posdef, nrhs = CASES[case][:3]
        n, ptr, row, val = gen()
        order = scal = None
        if len(CASES[case]) > 3 and CASES[case][3] == "matching":   # options%ordering = 2 + options%scaling = 3
            from spral_b200 import scaling as S
            order, scal, _ = S.match_order_metis(n, ptr, row, val)
        from oracle_engine import OracleEngine
        ctx = sdist.DistContext(world, rank, 0, engine=OracleEngine())
        ak = sdist.analyse(ctx, n, ptr, row, order=order)
        a = ak.analysis
        A = M.to_scipy(n, ptr, row, val)
        rng = np.random.default_rng(11)
        X = np.asfortranarray(rng.uniform(-1, 1, (n, nrhs)))
        B = np.asfortranarray(A @ X)
        out = {}
        for rep in range(2):                       # twice: epochs / store keys must not collide
            fk = sdist.factor(ctx, ak, posdef, val, scaling=scal)
            inform = sdist.reduce_inform(ctx, fk.inform)
            Xs = sdist.solve(ctx, fk, B)
            out = dict(bwd=float(ref.backward_error(A, Xs, B)), inform=inform, nparts=int(a.nparts),
                       owners=sorted(set(ak.rank_of)), x0=Xs[:5, 0].tolist())
            # forward then diag+backward compose to the full solve
            Y = sdist.solve(ctx, fk, sdist.solve(ctx, fk, B, job=1), job=4)
            out["compose"] = float(np.abs(Y - Xs).max())
            sdist.free(fk) if False else None
        # single-process reference on the same analysis
        if rank == 0:
            parts, r, sc = ref.ref_factor(a, posdef, val, scaling=scal)
            Xr = ref.ref_solve(a, parts, posdef, B, sc)
            out["ref_bwd"] = float(ref.backward_error(A, Xr, B))
            out["ref_inform"] = {k: int(r[k]) for k in ("num_neg", "num_two", "num_delay", "num_factor", "matrix_rank")}
            out["maxdiff"] = float(np.abs(Xr - Xs).max() / max(1.0, np.abs(Xr).max()))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception as e:          # pragma: no cover
        import traceback
        q.put((rank, dict(error=traceback.format_exc())))


def _cases():
    from spral_b200 import matrices as M
    return {
        "st27_indef": (lambda: M.stencil_3d_27pt(12, shift=13.0), False, 3),
        "lap3d_posdef": (lambda: M.laplacian_3d_7pt(12), True, 2),
        "kkt_delays": (lambda: M.kkt_saddle(3000), False, 1),
        "kkt_grid_matching": (lambda: M.kkt_grid(12), False, 2, "matching"),   # cfg4 in small, across ranks
    }


CASES = _cases()


@pytest.mark.parametrize("case,world", [("st27_indef", 2), ("lap3d_posdef", 2), ("kkt_delays", 2), ("st27_indef", 3),
                                        ("kkt_grid_matching", 2)])
def test_two_rank_factor_solve(case, world):
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r, out in res.items():
        assert "error" not in out, out.get("error")
    o = res[0]
    assert o["nparts"] > 1 and len(o["owners"]) == world      # the work really was split
    assert o["bwd"] < 5e-11 and o["bwd"] <= 20 * o["ref_bwd"] + 1e-15
    assert o["compose"] < 1e-9
    for k in ("num_neg", "num_delay", "num_factor", "matrix_rank"):   # same engine, same parts -> identical
        assert int(o["inform"][k]) == o["ref_inform"][k], (k, o["inform"][k], o["ref_inform"][k])
    for r in range(1, world):                                      # every rank returns the same solution
        np.testing.assert_allclose(res[r]["x0"], o["x0"], rtol=1e-12, atol=1e-14)


def test_assign_ranks_every_part_owned_and_roots_on_heaviest_child():
    from spral_b200 import matrices as M, dist as sdist
    from spral_b200.ssids import Analysis
    n, ptr, row, val = M.stencil_3d_27pt(16, shift=13.0)
    for world in (2, 4, 8):
        a = Analysis(n, ptr, row, ngpu=world)
        rank_of = sdist.assign_ranks(a, world)
        consumer, children = sdist.part_graph(a)
        assert len(rank_of) == a.nparts and all(0 <= r < world for r in rank_of)
        assert sum(1 for q in consumer if q < 0) >= 1
        for p in range(a.nparts):
            assert consumer[p] == -1 or consumer[p] > p           # postorder: consumers come later
            if int(a.exec_loc[p]) == -1 and children[p]:
                assert rank_of[p] in {rank_of[c] for c in children[p]}
        a.close()


def _worker_failure(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch.distributed as dist
        import oracle_ref as ref
        ref.ensure_env()
        from datetime import timedelta
        from spral_b200 import matrices as M, dist as sdist
        from oracle_engine import OracleEngine
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=timedelta(seconds=60))
        n, ptr, row, val = M.stencil_3d_27pt(12, shift=13.0)          # indefinite ...
        ctx = sdist.DistContext(world, rank, 0, engine=OracleEngine())
        ak = sdist.analyse(ctx, n, ptr, row)
        fk = sdist.factor(ctx, ak, True, val)                          # ... factorised as positive definite
        inform = sdist.reduce_inform(ctx, fk.inform)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, dict(flag=int(inform["flag"]), local_flag=int(fk.inform["flag"]), nparts=int(ak.analysis.nparts))))
    except Exception:          # pragma: no cover
        import traceback
        q.put((rank, dict(error=traceback.format_exc())))


def test_an_error_in_one_part_reaches_every_rank_without_hanging():
    """A part that ends with an error flag publishes that instead of its contribution block: the ranks
    that wait for it stop too (inform%flag = -6 everywhere), nobody sits in the rendezvous store."""
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = _free_port()
    world = 2
    procs = [ctxmp.Process(target=_worker_failure, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r, out in res.items():
        assert "error" not in out, out.get("error")
        assert out["flag"] == -6 and out["nparts"] > 1, out
