"""One process per GPU: the part-level scheduler of ssids_factor / ssids_solve.

Mirrors fkeep%inner_factor / inner_solve (src/ssids/fkeep.F90:61-323 of the
reference), where the reference runs one OpenMP thread per GPU inside one
process and hands contribution blocks through host memory, this driver runs
one PROCESS per GPU (torchrun) and

  * every part of the subtree partition (find_subtree_partition, anal.F90:289-464)
    is owned by exactly one rank; the parts the reference would leave to the CPU
    ("all-region" parts, exec_loc = -1) are given to the rank of their heaviest
    child, so the root of the tree is factorised on a GPU too;
  * a contribution block crosses ranks only where a subtree feeds an ancestor on
    another GPU: the producer publishes a CUDA IPC handle in the rendezvous
    store, the consumer pulls the block over NVLink (peer copy) when it reaches
    the parent part.  Nothing blocks the producer, so the schedule cannot
    deadlock: every rank walks its parts in increasing (post)order and only ever
    waits for a lower-numbered part;
  * the solves exchange, per cross-rank edge, the update a part makes to the
    rows of its ancestors (forward) and the ancestors' solution on those rows
    (backward); the final solution is one all-reduce.

The numeric engine is reached through a small adapter (`GpuEngine` below: symbolic
subtree, factor, solve, contribution hand-over).  The product knows no other engine; the
world_size-2 gloo tests on CPU inject their own adapter (tests/oracle_engine.py), so the
host logic of this file is exercised without a GPU and without any engine-specific code here.
"""
import ctypes as C
import os
import pickle
import sys
import time

import numpy as np

from . import _lib
from ._lib import Contrib, Options
from .ssids import Analysis, SymbolicSubtree, free_contrib


class GpuEngine:
    """The B200 engine behind the C ABI (the only engine of the product).  An adapter offers:
    device_ipc (contribution blocks stay on the device and cross ranks by CUDA IPC; otherwise the
    payload is staged through the rendezvous store), device(local_rank) for the right-hand sides,
    symbolic / factor / solve / get_contrib / device_ms."""
    device_ipc = True

    def device(self, local_rank):
        import torch
        return torch.device("cuda", local_rank)

    def symbolic(self, analysis, part, local_rank, options):
        return SymbolicSubtree(analysis, part, device=local_rank, options=options)

    @staticmethod
    def _join_torch_stream():
        """The engine runs on private non-blocking streams: whatever torch has queued on ITS current stream
        (index_select / scaling of the right-hand side, a device-resident value array being filled) must have
        finished before the engine reads it.  The return path is ordered by the engine's own synchronisation."""
        import torch
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()

    def factor(self, symb, analysis, part, posdef, val, child_contrib, options, scaling):
        self._join_torch_stream()
        return symb.factor(posdef, val, child_contrib, options, scaling)

    def solve(self, ns, which, X, nrhs, n):
        self._join_torch_stream()
        getattr(ns, f"solve_{which}")(X.data_ptr(), nrhs, n)

    def get_contrib(self, ns):
        return ns.get_contrib(device_resident=True)

    def device_ms(self, ns):
        return float(ns.timings()[1])


class DistContext:
    def __init__(self, world=1, rank=0, local_rank=0, engine=None, store=None, tag="ssids"):
        if engine is None or engine == "gpu":
            engine = GpuEngine()
        elif isinstance(engine, str):
            raise ValueError(f"unknown engine {engine!r}: the product has the GPU engine only; tests inject an adapter object")
        self.world, self.rank, self.local_rank, self.engine = world, rank, local_rank, engine
        self._store, self.tag, self.epoch = store, tag, 0
        self._stage = {}          # producer part -> (device block, capacity): re-used across factorisations
        self._exported = {}       # producer part -> epochs whose export buffers may still be read by the consumer

    @property
    def store(self):
        if self._store is None:
            import torch.distributed as dist
            self._store = dist.distributed_c10d._get_default_store()
        return self._store


# --------------------------------------------------------------------------
# ownership of parts
# --------------------------------------------------------------------------

def part_graph(a):
    """consumer[p] = part that receives p's contribution (or -1); children[q]."""
    nparts = a.nparts
    consumer = [-1] * nparts
    for p in range(nparts):
        idx = int(a.contrib_idx[p])
        if idx > nparts:
            continue
        for q in range(p + 1, nparts):
            if int(a.contrib_ptr[q]) <= idx < int(a.contrib_ptr[q + 1]):
                consumer[p] = q
                break
    children = [[] for _ in range(nparts)]
    for p, q in enumerate(consumer):
        if q >= 0:
            children[q].append(p)
    return consumer, children


def part_flops(a):
    fl = np.zeros(a.nparts)
    for p in range(a.nparts):
        for node in range(int(a.part[p]), int(a.part[p + 1])):
            m = int(a.rptr[node] - a.rptr[node - 1])
            nc = int(a.sptr[node] - a.sptr[node - 1])
            j = np.arange(nc, dtype=np.float64)
            fl[p] += float(((m - j) ** 2).sum())
    return fl


def assign_ranks(a, world):
    """rank_of[p].  Leaf parts: the GPU chosen by the analyse phase (exec_loc,
    anal.F90:496-501).  All-region parts (exec_loc = -1): the rank of the child
    whose subtree carries the most flops (it finishes last, so its block never
    has to travel)."""
    consumer, children = part_graph(a)
    fl = part_flops(a)
    sub = fl.copy()
    rank_of = [0] * a.nparts
    for p in range(a.nparts):
        loc = int(a.exec_loc[p])
        for c in children[p]:
            sub[p] += sub[c]
        if loc >= 2:
            rank_of[p] = (loc - 2) % world
        elif children[p]:
            rank_of[p] = rank_of[max(children[p], key=lambda c: sub[c])]
        else:
            rank_of[p] = 0
    return rank_of


def node_flops(a):
    """Factorisation flops per node without delays: sum_{j < ncol} (m - j)^2 (cpu/factor.hxx:121-124)."""
    m = (a.rptr[1:] - a.rptr[:-1]).astype(np.float64)
    nc = (a.sptr[1:] - a.sptr[:-1]).astype(np.float64)
    # sum_{j=0}^{nc-1} (m-j)^2 = nc m^2 - m nc (nc-1) + (nc-1) nc (2nc-1) / 6
    return nc * m * m - m * nc * (nc - 1) + (nc - 1) * nc * (2 * nc - 1) / 6.0


def proportional_partition(a, world, small=0.02, whole_ok=True):
    """Proportional mapping of the assembly tree onto `world` ranks (one process per GPU).

    The reference's find_subtree_partition (anal.F90:289-464) cuts subtrees for the GPUs and leaves EVERYTHING
    above the cut to one "all-region" part -- on the CPU in the reference, on one GPU here -- which for a 3-D
    problem is most of the work (cfg5: 60 % of the flops above an 8-way cut).  Here the ranks of a node are dealt to
    its children in proportion to their subtree flops; a subtree that ends up with one rank is a leaf part, and every
    node that still has several ranks below it becomes a part of its own, run on the rank of its heaviest child (whose
    contribution block then never travels).  So the level of 6 fronts below the top runs on 6 GPUs, the level of 3 on
    3, and only the root front is serial (and is the one the distributed top front splits further).
    Returns (part, rank_of): part = 1-based first nodes (len nparts + 1) in postorder."""
    nn = a.nnodes
    par = np.asarray(a.sparent, dtype=np.int64) - 1                # 0-based parent, nn for roots
    fl = node_flops(a)
    sub = fl.copy()
    first = np.arange(nn)                                          # first node of the subtree (postorder: contiguous)
    children = [[] for _ in range(nn + 1)]
    for i in range(nn):
        p = int(par[i]) if par[i] < nn else nn
        children[p].append(i)
        if p < nn:
            sub[p] += sub[i]
            first[p] = min(first[p], first[i])
    total = float(fl.sum()) or 1.0
    parts, ranks_out = [], []
    load = [0.0] * world                                           # flops dealt to every rank so far (small subtrees)

    def leaf(node, rank):
        parts.append(int(first[node]) + 1)
        ranks_out.append(rank)
        load[rank] += float(sub[node])

    def assign(node, ranks):
        """Emits the parts of the subtree of `node` in postorder; returns the rank of the part that holds `node`."""
        if len(ranks) == 1 or not children[node]:
            leaf(node, ranks[0])
            return ranks[0]
        ch = children[node]                                        # increasing index == postorder
        big = [c for c in ch if sub[c] >= small * total / world * len(ranks) and sub[c] > 0]
        if not big:
            leaf(node, ranks[0])
            return ranks[0]
        if len(big) == 1:
            # a chain: follow it down to the first node that branches (two or more big children).  The subtree of that
            # node is mapped onto all the ranks; the small subtrees that hang off the chain BEFORE it in postorder are
            # leaf parts, everything after it -- the other small subtrees and the chain nodes themselves -- is one
            # part (contiguous, its only exit is `node`).  KKT matrices have such chains near the top of the tree.
            chain, x = [node], big[0]
            while True:
                chx = children[x]
                bx = [c for c in chx if sub[c] >= small * total / world * len(ranks) and sub[c] > 0]
                if len(bx) != 1:
                    break
                chain.append(x)
                x = bx[0]
            if len(children[x]) < 2 or sub[x] < 0.5 * sub[node]:   # nothing worth spreading below the chain
                leaf(node, ranks[0])
                return ranks[0]
            before = []                                            # small subtrees with indices below first[x]
            for cn in chain:
                for c in children[cn]:
                    if c != x and c not in chain and c < first[x]:
                        before.append(c)
            for c in sorted(before):
                leaf(c, min(ranks, key=lambda q: load[q]))
            r = assign(x, ranks)
            parts.append(x + 2)                                    # nodes x+1 .. node (1-based first node: x + 2)
            ranks_out.append(r)
            load[r] += float(sub[node] - sub[x] - sum(sub[c] for c in before))
            return r
        alloc = {}
        if len(big) > len(ranks):
            # more big children than ranks: longest-processing-time packing, one rank per bin -- except for a child
            # that outweighs a fair share of this node's ranks: it is dealt ALL of them (its own children are then
            # spread with the loads of its smaller siblings already on the books)
            order = sorted(big, key=lambda c: -sub[c])
            fair = (sum(float(sub[c]) for c in big) + float(fl[node]) + sum(load[r] for r in ranks)) / len(ranks)
            trial = {r: load[r] for r in ranks}
            for i, c in enumerate(order):                          # what plain packing would give (the node itself goes
                r = min(ranks, key=lambda q: trial[q])             # where its heaviest child is)
                trial[r] += float(sub[c]) + (float(fl[node]) if i == 0 else 0.0)
            whole = order[0] if whole_ok and max(trial.values()) > 1.15 * fair and len(children[order[0]]) > 1 else None
            bins = {r: load[r] for r in ranks}
            r0 = min(ranks, key=lambda q: bins[q])                 # takes the chain: this node and the top of `whole`
            if whole is not None:
                bins[r0] += float(fl[node]) + float(fl[whole])
            for c in order:
                if c == whole:
                    continue
                r = min(ranks, key=lambda q: bins[q])
                alloc[c] = [r]
                bins[r] += float(sub[c])
            if whole is not None:
                alloc[whole] = [r0] + sorted([q for q in ranks if q != r0], key=lambda q: bins[q])
        else:
            # integer shares of the ranks for the big children, proportional to their flops, at least one each
            k = {c: 1 for c in big}
            for _ in range(len(ranks) - len(big)):
                c = max(big, key=lambda c: sub[c] / k[c])
                k[c] += 1
            pos = 0
            for c in sorted(big, key=lambda c: -sub[c]):
                alloc[c] = ranks[pos:pos + k[c]]
                pos += k[c]
        top_rank, heaviest = ranks[0], -1.0
        prebooked = {}
        for c in ch:                                               # loads of the single-rank siblings count before anybody recurses
            if c in alloc and len(alloc[c]) == 1:
                prebooked[c] = float(sub[c])
                load[alloc[c][0]] += prebooked[c]
        for c in ch:
            if c in alloc:
                if c in prebooked:
                    load[alloc[c][0]] -= prebooked[c]                  # leaf() books it again
                r = assign(c, alloc[c])
            else:                                                  # small subtree (or no rank left): least loaded rank of this node
                r = min(ranks, key=lambda q: load[q])
                leaf(c, r)
            if sub[c] > heaviest:
                heaviest, top_rank = sub[c], r
        parts.append(node + 1)
        ranks_out.append(top_rank)
        load[top_rank] += float(fl[node])
        return top_rank

    roots = children[nn]
    if len(roots) == 1:
        assign(roots[0], list(range(world)))
    else:                                                          # a forest: deal the ranks to the trees like children of a virtual root
        order = sorted(roots, key=lambda c: -sub[c])
        shares = {c: [] for c in roots}
        for i, r in enumerate(range(world)):
            shares[order[i % len(order)] if i < len(order) else max(order, key=lambda c: sub[c] / max(1, len(shares[c])))].append(r)
        for c in roots:
            if shares[c]:
                assign(c, shares[c])
            else:
                leaf(c, min(range(world), key=lambda q: load[q]))
    parts.append(nn + 1)
    return np.asarray(parts, dtype=np.int32), ranks_out


class DistAkeep:
    def __init__(self, analysis, subtrees, rank_of, consumer, children):
        self.analysis, self.subtrees = analysis, subtrees
        self.rank_of, self.consumer, self.children = rank_of, consumer, children
        # list-scheduling priority of a part: flops on the path from it to the root of the tree ("bottom level").
        # A rank works through its parts in this order, so what another rank is waiting for comes first
        # (children always precede their consumers: their bottom level is larger).
        fl = part_flops(analysis)
        self.priority = [0.0] * analysis.nparts
        for p in range(analysis.nparts - 1, -1, -1):
            q = consumer[p]
            self.priority[p] = float(fl[p]) + (self.priority[q] if q >= 0 else 0.0)


def modelled_critical_path(a, world, rank_of=None):
    """Length of the schedule when every part costs its flops (normalised to the whole tree) plus a fixed
    per-part overhead and every rank works through its parts by decreasing bottom level (DistAkeep.priority)."""
    consumer, children = part_graph(a)
    if rank_of is None:
        rank_of = assign_ranks(a, world)
    fl = part_flops(a)
    tot = max(float(fl.sum()), 1.0)
    prio = [0.0] * a.nparts
    for p in range(a.nparts - 1, -1, -1):
        prio[p] = float(fl[p]) + (prio[consumer[p]] if consumer[p] >= 0 else 0.0)
    fin = [0.0] * a.nparts
    busy = [0.0] * world
    for p in sorted(range(a.nparts), key=lambda p: (-prio[p], p)):
        start = max([fin[c] for c in children[p]] + [busy[rank_of[p]]])
        fin[p] = start + fl[p] / tot + 0.005
        busy[rank_of[p]] = fin[p]
    return max(fin) if fin else 0.0


def analyse(ctx, n, ptr, row, order=None, nemin=32, options=None, tune_partition=True, **kw):
    """ssids_analyse on every rank (deterministic, replicated), symbolic subtrees
    only for the parts this rank owns.  With several GPUs the subtree partition is
    computed for a few values of options%max_load_inbalance (the reference's own
    knob, src/ssids/datatypes.f90:223-228) and the one with the shortest modelled
    critical path is kept: a finer partition is not always better, every part
    boundary costs the level-set batching across its subtrees."""
    a = Analysis(n, ptr, row, order=order, nemin=nemin, ngpu=ctx.world, **kw)
    mode = os.environ.get("SPRAL_B200_PARTITION", "proportional")
    if ctx.world > 1 and tune_partition and mode == "proportional" and a.nnodes > 0:
        best = None
        for whole_ok in (False, True):                           # two variants of the mapping: the schedule model picks
            part, ranks = proportional_partition(a, ctx.world, whole_ok=whole_ok)
            a.set_partition(part, [r + 2 for r in ranks])
            cp = modelled_critical_path(a, ctx.world, list(ranks))
            if best is None or cp < best[0] - 1e-9:
                best = (cp, part, ranks)
        _, part, ranks = best
        a.set_partition(part, [r + 2 for r in ranks])
        consumer, children = part_graph(a)
        subtrees = [ctx.engine.symbolic(a, p, ctx.local_rank, options) if ranks[p] == ctx.rank else None
                    for p in range(a.nparts)]
        return DistAkeep(a, subtrees, list(ranks), consumer, children)
    if ctx.world > 1 and tune_partition and "max_load_inbalance" not in kw:
        best, best_cp = a, modelled_critical_path(a, ctx.world)
        for mli in (2.0, 3.0):
            kw2 = dict(kw, max_load_inbalance=mli)
            cand = Analysis(n, ptr, row, order=a.order.copy(), nemin=nemin, ngpu=ctx.world, **kw2)
            cp = modelled_critical_path(cand, ctx.world)
            if cand.nparts > 1 and cp < best_cp - 1e-9:
                best.close() if best is not a else None
                best, best_cp = cand, cp
            else:
                cand.close()
        if best is not a:
            a.close()
        a = best
    rank_of = assign_ranks(a, ctx.world)
    consumer, children = part_graph(a)
    subtrees = []
    for p in range(a.nparts):
        if rank_of[p] != ctx.rank:
            subtrees.append(None)
        else:
            subtrees.append(ctx.engine.symbolic(a, p, ctx.local_rank, options))
    return DistAkeep(a, subtrees, rank_of, consumer, children)


# --------------------------------------------------------------------------
# transport of contribution blocks
# --------------------------------------------------------------------------

def _key(ctx, kind, p):
    return f"{ctx.tag}/{ctx.epoch}/{kind}/{p}"


def _delete(ctx, key):
    """Spent rendezvous keys are removed (a long-running solve loop must not grow the store)."""
    try:
        ctx.store.delete_key(key)
    except Exception:           # a store without delete_key (FileStore on some builds): keys are small, keep them
        pass


def _contrib_rlist(a, p):
    """Global row indices of part p's root contribution (host, from the replicated analysis)."""
    root = int(a.part[p + 1]) - 1
    lo = int(a.rptr[root - 1]) - 1 + int(a.sptr[root] - a.sptr[root - 1])
    hi = int(a.rptr[root]) - 1
    return np.ascontiguousarray(a.rlist[lo:hi], dtype=np.int32)


def publish_contrib(ctx, ak, p, ns):
    """Producer side of a cross-rank edge."""
    if ctx.engine.device_ipc:
        lib = _lib.load()
        # the engine packs into one of TWO export buffers per part, alternating: the buffer of two exports ago is
        # re-used now, so its consumer must have pulled it (it acknowledges with a "pulled" key, deleted here)
        hist = ctx._exported.setdefault(p, [])
        while len(hist) >= 2:
            old = _Epoch(ctx, hist.pop(0))
            ctx.store.wait([_key(old, "pulled", p)])
            _delete(ctx, _key(old, "pulled", p))
        hist.append(ctx.epoch)
        handle = (C.c_ubyte * 64)()
        n, nd, nbytes, blk = C.c_int(), C.c_int(), C.c_int64(), C.c_void_p()
        rc = lib.spral_ssids_gpu_subtree_export_contrib_ipc(ns._h, handle, C.byref(n), C.byref(nd),
                                                            C.byref(nbytes), C.byref(blk))
        if rc != 0:
            raise RuntimeError(f"export_contrib_ipc failed: cudaError {rc}")
        meta = dict(kind="ipc", handle=bytes(handle), n=n.value, ndelay=nd.value, bytes=nbytes.value)
    else:
        c = ns.get_contrib()
        n, nd = c.n, c.ndelay

        def arr(ptr, count, ctype):
            if not ptr or count == 0:
                return None
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), (count,)).copy()
        val = arr(c.val, c.ldval * n, C.c_double)
        rows = nd + n
        dval = arr(c.delay_val, c.lddelay * nd, C.c_double)
        meta = dict(kind="host", n=n, ndelay=nd, ldval=c.ldval, lddelay=c.lddelay, val=val, dval=dval,
                    dperm=arr(c.delay_perm, nd, C.c_int))
    ctx.store.set(_key(ctx, "contrib", p), pickle.dumps(meta))
    return meta


class _Fetched:
    """Keeps the memory behind a fetched Contrib alive / frees the device block."""

    def __init__(self, contrib, keep, block=None):
        self.contrib, self.keep, self.block = contrib, keep, block

    def release(self):
        if self.block:
            _lib.load().spral_ssids_b200_device_free(self.block)
            self.block = None


class PartFailed(Exception):
    """A part this one depends on ended with an error flag on another rank."""

    def __init__(self, flag):
        super().__init__(f"a child part failed with flag {flag}")
        self.flag = flag


def publish_failure(ctx, ak, p, flag):
    """Tells the (remote) consumer of part p that no contribution will come: without it that rank
    would wait in fetch_contrib until the store times out."""
    q = ak.consumer[p]
    if q >= 0 and ak.rank_of[q] != ctx.rank:
        ctx.store.set(_key(ctx, "contrib", p), pickle.dumps(dict(kind="failed", flag=int(flag))))


def fetch_contrib(ctx, ak, p, posdef):
    """Consumer side: waits for part p's block and brings it to this rank."""
    key = _key(ctx, "contrib", p)
    ctx.store.wait([key])
    meta = pickle.loads(ctx.store.get(key))
    if meta["kind"] == "failed":
        raise PartFailed(meta["flag"])
    a = ak.analysis
    rl = _contrib_rlist(a, p)
    n, nd = meta["n"], meta["ndelay"]
    c = Contrib()
    c.n, c.ndelay, c.posdef, c.owner, c.owner_ptr, c.ready = n, nd, posdef, 2, None, 1
    c.rlist = rl.ctypes.data
    if meta["kind"] == "ipc":
        lib = _lib.load()
        blk, cap = ctx._stage.get(p, (None, 0))
        if cap < meta["bytes"]:
            if blk:
                lib.spral_ssids_b200_device_free(blk)
            cap = int(meta["bytes"] * 1.1) + 256
            blk = lib.spral_ssids_b200_device_alloc(ctx.local_rank, cap)
            if not blk:
                raise MemoryError("device_alloc failed")
            ctx._stage[p] = (blk, cap)
        handle = (C.c_ubyte * 64).from_buffer_copy(meta["handle"])
        rc = lib.spral_ssids_b200_ipc_pull(ctx.local_rank, handle, meta["bytes"], blk)
        if rc != 0:
            raise RuntimeError(f"ipc_pull failed: cudaError {rc}")
        _delete(ctx, key)                                   # single consumer: the key is spent
        ctx.store.set(_key(ctx, "pulled", p), b"1")         # the producer may re-use that export buffer
        b_val = n * n * 8
        rows = nd + n
        c.val, c.ldval = (blk if n else None), n
        c.delay_val = blk + b_val if nd else None
        c.lddelay = rows
        c.delay_perm = blk + b_val + rows * nd * 8 if nd else None
        c.device = ctx.local_rank
        return _Fetched(c, [rl], None)          # the staging block stays with the context
    _delete(ctx, key)
    keep = [rl, meta["val"], meta["dval"], meta["dperm"]]
    c.val = meta["val"].ctypes.data if meta["val"] is not None else None
    c.ldval = meta["ldval"]
    c.delay_val = meta["dval"].ctypes.data if meta["dval"] is not None else None
    c.lddelay = meta["lddelay"]
    c.delay_perm = meta["dperm"].ctypes.data if meta["dperm"] is not None else None
    c.device = -1
    return _Fetched(c, keep)


# --------------------------------------------------------------------------
# factor
# --------------------------------------------------------------------------

class DistFkeep:
    def __init__(self, ak, posdef, numeric, inform, ext_rows, scaling, epoch):
        self.akeep, self.posdef, self.numeric, self.inform = ak, posdef, numeric, inform
        self.ext_rows, self.scaling, self.epoch = ext_rows, scaling, epoch


def _new_inform(a):
    return dict(flag=0, num_delay=0, num_factor=0, num_flops=0, num_neg=0, num_two=0, maxfront=0,
                maxsupernode=0, matrix_rank=0, num_zero=0, not_first_pass=0, not_second_pass=0, cuda_error=0)


def _accumulate(inform, st):
    """cpu_copy_stats_out (src/ssids/cpu/cpu_iface.f90:74-94)."""
    if st.flag < 0:
        inform["flag"] = min(inform["flag"], st.flag) if inform["flag"] < 0 else st.flag
        inform["cuda_error"] = getattr(st, "cuda_error", 0)
        return
    if inform["flag"] >= 0:
        inform["flag"] = max(inform["flag"], st.flag)
    for k in ("num_delay", "num_factor", "num_flops", "num_neg", "num_two", "num_zero",
              "not_first_pass", "not_second_pass"):
        inform[k] += getattr(st, k)
    inform["maxfront"] = max(inform["maxfront"], st.maxfront)
    inform["maxsupernode"] = max(inform["maxsupernode"], st.maxsupernode)


def factor(ctx, ak, posdef, val, options=None, scaling=None):
    """fkeep%inner_factor: this rank's parts in postorder; `val` is a numpy array or
    a raw (host or device) pointer to the values of A."""
    a = ak.analysis
    ctx.epoch += 1
    sc = None
    if scaling is not None:
        sc = np.ascontiguousarray(np.asarray(scaling, dtype=np.float64)[a.invp - 1])
    nparts = a.nparts
    local = {}                        # part -> Contrib produced on this rank (device resident)
    numeric = [None] * nparts
    ext_rows = [None] * nparts        # rows of x outside the part that it touches (solve exchange)
    inform = _new_inform(a)
    trace = os.environ.get("SPRAL_B200_TRACE")
    t_start = time.perf_counter()
    import threading
    from concurrent.futures import ThreadPoolExecutor
    lock = threading.Lock()
    futures = {}
    failed = threading.Event()

    def run_part(p):
        """One part: wait for the local children, pull the remote ones, factorise, hand on."""
        for c_part in ak.children[p]:
            if ak.rank_of[c_part] == ctx.rank:
                futures[c_part].result()
        if failed.is_set():
            publish_failure(ctx, ak, p, inform["flag"] if inform["flag"] < 0 else -99)
            return
        t_p0 = time.perf_counter()
        cc, fetched = [], []
        try:
            for c_part in ak.children[p]:
                if ak.rank_of[c_part] == ctx.rank:
                    cc.append(local.pop(c_part))
                else:
                    f = fetch_contrib(ctx, ak, c_part, posdef)
                    fetched.append(f)
                    cc.append(f.contrib)
        except PartFailed as e:                       # the error of another rank's part ends this one too
            with lock:
                inform["flag"] = min(inform["flag"], e.flag) if inform["flag"] < 0 else e.flag
            failed.set()
            for f in fetched:
                f.release()
            publish_failure(ctx, ak, p, e.flag)
            return
        # children[] is ordered by part; the slots contrib_ptr[p].. follow the same order
        t_p1 = time.perf_counter()
        ns = ctx.engine.factor(ak.subtrees[p], a, p, posdef, val, cc, options, sc)
        st = ns.stats
        numeric[p] = ns
        for c in cc:
            if c.owner == 1:
                free_contrib(c)
        for f in fetched:
            f.release()
        with lock:
            _accumulate(inform, st)
        if st.flag < 0:
            failed.set()
            publish_failure(ctx, ak, p, st.flag)
            return
        t_p2 = time.perf_counter()
        q = ak.consumer[p]
        if q >= 0:
            c = ctx.engine.get_contrib(ns)
            rows = [_contrib_rlist(a, p)]
            if c.ndelay:
                if c.device >= 0:
                    dp = np.empty(c.ndelay, dtype=np.int32)
                    rc = _lib.load().spral_ssids_b200_copy_to_host(dp.ctypes.data, c.delay_perm, 4 * c.ndelay)
                    if rc != 0:
                        raise RuntimeError(f"copy_to_host failed: cudaError {rc}")
                else:
                    dp = np.ctypeslib.as_array(C.cast(c.delay_perm, C.POINTER(C.c_int)), (c.ndelay,)).copy()
                rows.append(dp)
            ext_rows[p] = np.concatenate(rows).astype(np.int64) - 1
            if ak.rank_of[q] == ctx.rank:
                local[p] = c
            else:
                publish_contrib(ctx, ak, p, ns)
                ctx.store.set(_key(ctx, "ext_rows", p), pickle.dumps(ext_rows[p]))
        if trace:
            t_p3 = time.perf_counter()
            print(f"[trace r{ctx.rank} e{ctx.epoch}] part {p}: start +{1e3*(t_p0-t_start):.1f} ms, fetch {1e3*(t_p1-t_p0):.1f}, "
                  f"factor {1e3*(t_p2-t_p1):.1f} (dev {ctx.engine.device_ms(ns):.1f}), "
                  f"publish {1e3*(t_p3-t_p2):.1f} ms, flops {st.num_flops:.3g}", file=sys.stderr, flush=True)

    mine = sorted((p for p in range(nparts) if ak.rank_of[p] == ctx.rank), key=lambda p: (-ak.priority[p], p))
    split = _split_roles(ctx, ak)                # distributed top front: (shm name, owner rank, helper ranks) or None
    if split and ctx.rank == split[1]:
        _lib.load().spral_ssids_b200_split_enable(ak.subtrees[nparts - 1]._h, split[0].encode(), len(split[2]))
    nthreads = max(1, int(os.environ.get("SPRAL_B200_PART_THREADS", "2")))
    if not ctx.engine.device_ipc or len(mine) <= 1:
        nthreads = 1
    # independent parts of this rank run concurrently (each on its own stream); tasks are
    # submitted by decreasing bottom level, so a parent never starts before its children have started
    with ThreadPoolExecutor(max_workers=nthreads) as ex:
        for p in mine:
            futures[p] = ex.submit(run_part, p)
        for p in mine:
            futures[p].result()
    if split and ctx.rank in split[2]:
        # this rank's parts are done: serve the owner of the top fronts until its part is finished
        rc = _lib.load().spral_ssids_b200_split_helper_serve(split[0].encode(), ctx.local_rank,
                                                             float(os.environ.get("SPRAL_B200_SPLIT_TIMEOUT", "20")),
                                                             split[2].index(ctx.rank))
        if trace:
            print(f"[trace r{ctx.rank} e{ctx.epoch}] split helper returned {rc}", file=sys.stderr, flush=True)
    return DistFkeep(ak, posdef, numeric, finish_inform(a, inform), ext_rows, sc, ctx.epoch)


def _split_roles(ctx, ak):
    """Distributed top front (csrc/split_front.h; SPRAL_B200_SPLIT=0 turns it off): the rank that owns the last part
    (the top of the tree) shares the trailing updates of its large fronts with up to SPRAL_B200_SPLIT_HELPERS other ranks
    (default 3: every helper receives every panel, so the owner's NVLink egress grows with their number -- measured on
    cfg5: 7 helpers are slower than 3); the far blocks are dealt round robin over the owner and the helpers.
    Returns (shared-memory name, owner, [helper ranks]) or None."""
    if os.environ.get("SPRAL_B200_SPLIT", "1") != "1" or ctx.world < 2 or not ctx.engine.device_ipc:
        return None
    lib = _lib.load()
    if not hasattr(lib, "spral_ssids_b200_split_helper_serve"):
        return None
    lib.spral_ssids_b200_split_enable.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.spral_ssids_b200_split_enable.restype = None
    lib.spral_ssids_b200_split_helper_serve.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int]
    lib.spral_ssids_b200_split_helper_serve.restype = C.c_int
    owner = ak.rank_of[ak.analysis.nparts - 1]
    nh = max(1, min(ctx.world - 1, int(os.environ.get("SPRAL_B200_SPLIT_HELPERS", "3"))))
    helpers = [(owner + 1 + i) % ctx.world for i in range(nh)]
    name = f"/spral_b200_split_{os.environ.get('MASTER_PORT', '0')}_{ctx.epoch}"
    return name, owner, helpers


def reduce_inform(ctx, inform):
    """inform%reduce over ranks (src/ssids/inform.f90:180-215)."""
    a_rank = None
    if ctx.world == 1:
        out = dict(inform)
    else:
        import torch
        import torch.distributed as dist
        dev = ctx.engine.device(ctx.local_rank)
        keys_sum = ("num_delay", "num_factor", "num_flops", "num_neg", "num_two", "num_zero",
                    "not_first_pass", "not_second_pass")
        t = torch.tensor([inform[k] for k in keys_sum], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out = dict(inform)
        for k, v in zip(keys_sum, t.tolist()):
            out[k] = int(v)
        t = torch.tensor([inform["maxfront"], inform["maxsupernode"], inform["flag"], -inform["flag"]],
                         dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mx = t.tolist()
        out["maxfront"], out["maxsupernode"] = int(mx[0]), int(mx[1])
        out["flag"] = int(-mx[3]) if mx[3] > 0 else int(mx[2])       # any error wins, else largest warning
    out["matrix_rank"] = inform.get("n_vars", 0) - out["num_zero"] if "n_vars" in inform else out.get("matrix_rank", 0)
    return out


def finish_inform(a, inform):
    inform = dict(inform)
    inform["n_vars"] = int(a.sptr[a.nnodes]) - 1
    inform["matrix_rank"] = inform["n_vars"] - inform["num_zero"]     # this rank's view; reduce_inform() completes it
    return inform


def free(fk):
    for ns in fk.numeric:
        if ns is not None:
            ns.close()
    fk.numeric = [None] * len(fk.numeric)


# --------------------------------------------------------------------------
# solve
# --------------------------------------------------------------------------

def _engine_solve(ctx, ns, which, X, nrhs, n):
    """X: torch tensor (nrhs, n) contiguous == column-major n x nrhs."""
    ctx.engine.solve(ns, which, X, nrhs, n)


def _send(ctx, kind, p, t):
    ctx.store.set(_key(ctx, kind, p), pickle.dumps(t.cpu().numpy()))


def _recv(ctx, kind, p, dev):
    import torch
    key = _key(ctx, kind, p)
    ctx.store.wait([key])
    t = torch.from_numpy(pickle.loads(ctx.store.get(key))).to(dev)
    _delete(ctx, key)                                       # one consumer per cross-rank edge
    return t


def solve(ctx, fk, x, job=0):
    """ssids_solve / inner_solve_cpu (src/ssids/fkeep.F90:234-323) across ranks.
    x: (n,) or (n, nrhs) numpy, same on every rank; returns the solution on every rank.
    The permutation to pivot order and the scaling (fkeep.F90:252-266, 300-315) run on
    the device that holds the factors."""
    import torch
    a = fk.akeep.analysis
    n = a.n
    dev = ctx.engine.device(ctx.local_rank)
    x = np.asarray(x, dtype=np.float64)
    one = x.ndim == 1
    Xh = x.reshape(n, -1)
    X = torch.from_numpy(np.ascontiguousarray(Xh.T)).to(dev)         # (nrhs, n), original order
    out = solve_device(ctx, fk, X, job)
    res = np.asfortranarray(out.cpu().numpy().T)
    return res[:, 0] if one else res


def solve_device(ctx, fk, X, job=0):
    """Same with X a torch tensor (nrhs, n) in ORIGINAL variable order on the compute
    device; returns a tensor of the same shape (device resident end to end)."""
    import torch
    a = fk.akeep.analysis
    dev = X.device
    cache = getattr(fk, "_perm_cache", None)
    if cache is None or cache[0].device != dev:
        invp = torch.from_numpy(np.asarray(a.invp, dtype=np.int64) - 1).to(dev)
        sc = torch.from_numpy(fk.scaling).to(dev) if fk.scaling is not None else None
        cache = fk._perm_cache = (invp, sc)
    invp, sc = cache
    X2 = X.index_select(1, invp).contiguous()                       # x2(i) = x(invp(i))
    if sc is not None and job in (0, 1):
        X2 *= sc[None, :]
    X2 = solve_pivot_order(ctx, fk, X2, job)
    if sc is not None and job in (0, 3, 4):
        X2 = X2 * sc[None, :]
    out = torch.empty_like(X2)
    out.index_copy_(1, invp, X2)                                    # x(invp(i)) = x2(i)
    return out


def solve_pivot_order(ctx, fk, X, job=0):
    """The part-level sweeps on a resident right-hand side.  X: torch tensor of
    shape (nrhs, n), contiguous (== column-major n x nrhs), in pivot order and
    already scaled; on the GPU engine it lives in HBM and never leaves it."""
    import torch
    ak = fk.akeep
    a = ak.analysis
    n = a.n
    nrhs = X.shape[0]
    dev = X.device
    ctx.epoch += 1
    nparts = a.nparts
    mine = [p for p in range(nparts) if ak.rank_of[p] == ctx.rank]
    ext = {}
    fkeys = _Epoch(ctx, fk.epoch)

    def rows_of(p):
        if p not in ext:
            r = fk.ext_rows[p]
            if r is None:                                        # produced on another rank
                key = _key(fkeys, "ext_rows", p)
                ctx.store.wait([key])
                r = pickle.loads(ctx.store.get(key))
                fk.ext_rows[p] = r
            ext[p] = torch.from_numpy(np.asarray(r, dtype=np.int64)).to(dev)
        return ext[p]

    delta = {}
    if job in (0, 1):                                            # forward
        for p in mine:
            q = ak.consumer[p]
            E = rows_of(p) if q >= 0 else None
            if E is not None:
                saved = X[:, E].clone()
                X[:, E] = 0.0
            for c in ak.children[p]:
                d = delta.pop(c) if ak.rank_of[c] == ctx.rank else _recv(ctx, "fwd", c, dev)
                X[:, rows_of(c)] += d
            _engine_solve(ctx, fk.numeric[p], "fwd", X, nrhs, n)
            if E is not None:
                d = X[:, E].clone()
                X[:, E] = saved
                if ak.rank_of[q] == ctx.rank:
                    delta[p] = d
                else:
                    _send(ctx, "fwd", p, d)
    if job == 2:
        for p in mine:
            _engine_solve(ctx, fk.numeric[p], "diag", X, nrhs, n)
    if job in (0, 3, 4):                                         # backward
        vals = {}
        which = "bwd" if job == 3 else "diag_bwd"
        for p in reversed(mine):
            q = ak.consumer[p]
            if q >= 0:
                E = rows_of(p)
                v = vals.pop(p) if ak.rank_of[q] == ctx.rank else _recv(ctx, "bwd", p, dev)
                saved = X[:, E].clone()
                X[:, E] = v
            _engine_solve(ctx, fk.numeric[p], which, X, nrhs, n)
            for c in ak.children[p]:
                v = X[:, rows_of(c)].clone()
                if ak.rank_of[c] == ctx.rank:
                    vals[c] = v
                else:
                    _send(ctx, "bwd", c, v)
            if q >= 0:
                X[:, E] = saved
    # every rank holds the final values of the variables eliminated in its own parts
    if ctx.world > 1:
        import torch.distributed as dist
        mask = torch.zeros(n, dtype=torch.float64, device=dev)
        for p in mine:                                           # increasing order
            lo = int(a.sptr[int(a.part[p]) - 1]) - 1
            hi = int(a.sptr[int(a.part[p + 1]) - 1]) - 1
            mask[lo:hi] = 1.0
            for c in ak.children[p]:                             # delays received are eliminated here
                mask[rows_of(c)[_ncontrib(a, c):]] = 1.0
            if ak.consumer[p] >= 0:                              # delays handed on are not
                mask[rows_of(p)[_ncontrib(a, p):]] = 0.0
        X = X * mask[None, :]
        dist.all_reduce(X, op=dist.ReduceOp.SUM)
    return X


def _ncontrib(a, p):
    root = int(a.part[p + 1]) - 1
    return int(a.rptr[root] - a.rptr[root - 1]) - int(a.sptr[root] - a.sptr[root - 1])


class _Epoch:
    """A view of the context pinned to another epoch (keys of the factor call)."""

    def __init__(self, ctx, epoch):
        self.tag, self.epoch = ctx.tag, epoch
