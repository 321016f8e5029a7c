#!/usr/bin/env python
"""bench.py -- ssids_factor FP64 GFLOP/s of the B200-native SSIDS engine.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            this engine
  python bench.py --impl reference ...                     the reference's own CPU
                                                            engine (oracle/_ref) on host cores
A "step" is one numeric factorisation (ssids_factor: fkeep%inner_factor over all
parts) of the BASELINE workload: 3-D 27-point stencil 100^3, shifted indefinite
(n = 10^6), LDL^T with threshold pivoting u = 0.01, METIS ordering, nemin = 32.
Analyse (ordering + symbolic) is set-up and is not timed, as in the metric.

value  = num_flops / device time of the factor call with A already resident in HBM
e2e    = num_flops / wall time of the same call through the C ABI with the HOST
         value array (pinned) -- H2D of A and D2H of the per-front results inside
roofline = the Schur-complement DMMA kernel (k_update, UPD_CONTRIB launches):
         algorithmic flops (m-n)(m-n+1)*nelim per front / CUDA-event time of those
         launches, against the FP64 tensor-pipe peak measured live on this GPU
The reference arm (--impl reference) and the cpu_baseline leg factorise THE SAME workload
(same matrix, ordering and options) with the reference's own CPU engine on all host cores,
and time its solves (1 and nrhs right-hand sides) beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("OMP_CANCELLATION", "TRUE")
os.environ.setdefault("NCCL_DEBUG", "WARN")       # keeps NCCL's version banner off stdout (one JSON line only)
# OMP_PROC_BIND belongs to the CPU reference arm only (run_reference): with it set, libgomp pins
# the initial thread of EVERY process that loads it (torch does) to the first CPU of the mask, and
# every thread created later inherits that mask -- all ranks of a torchrun job then share core 0.
_FULL_AFFINITY = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None

import numpy as np  # noqa: E402

SHIFT = 13.0
REF_MAX_STEPS = 2         # the reference arm times at most this many full factorisations (one takes 10-50 s of CPU)


def make_matrix(grid):
    from spral_b200 import matrices as M
    return M.stencil_3d_27pt(grid, shift=SHIFT)


def make_workload(args):
    """(description, posdef, scaling method or None, (n, ptr, row, val)).  The default -- what the driver
    measures -- is BASELINE configs[4] (cfg5); the others are the remaining BASELINE configs, for the
    developer's own runs (`--workload cfg2|cfg3|cfg4`)."""
    from spral_b200 import matrices as M
    w = args.workload
    if w == "cfg5":
        g = args.grid
        return (f"ssids_factor 3-D 27-point {g}^3 shifted (sigma={SHIFT}) indefinite, n={g ** 3}, LDL^T u=0.01, "
                f"METIS order, nemin=32 (BASELINE configs[4])", False, None, make_matrix(g))
    if w == "cfg3":
        g = args.grid if args.grid != 100 else 80
        return (f"ssids_factor 3-D 27-point {g}^3 shifted (sigma={SHIFT}) indefinite, LDL^T u=0.01 (BASELINE configs[2])",
                False, None, make_matrix(g))
    if w == "cfg2":
        g = args.grid if args.grid != 100 else 60
        return (f"ssids_factor 3-D 7-point Laplacian {g}^3, positive definite LL^T (BASELINE configs[1])", True, None,
                M.laplacian_3d_7pt(g))
    if w == "cfg4":
        g = args.grid if args.grid != 100 else 70
        return (f"ssids_factor structured KKT saddle point kkt_grid({g}), n={int(round(g ** 3 / 0.7))}, 30% zero-diagonal "
                f"rows, matching-based scaling (BASELINE configs[3])", False, "hungarian", M.kkt_grid(g))
    raise SystemExit(f"unknown workload {w}")


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons while the timed region runs (NVML in
    process -- spawning nvidia-smi every 200 ms measurably perturbs the run)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.max_mhz, self.reasons = index, [], 0, set()
        self._stop_evt = threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                pass
        return index

    def _sample(self):
        nv = self.nv
        if nv is not None:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                              ("hw_thermal_slowdown", 0x40)):
                if r & bit:
                    self.reasons.add(name)
        else:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                  str(self.index)], capture_output=True, text=True, timeout=5).stdout
            r = [c.strip() for c in out.strip().split(",")]
            self.sm.append(float(r[0])); self.max_mhz = max(self.max_mhz, float(r[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop_evt.wait(0.1 if self.nv is not None else 1.0)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": float(self.max_mhz) or None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def host_cores():
    """CPUs this process may run on (cgroup cpusets / taskset respected)."""
    return len(_FULL_AFFINITY) if _FULL_AFFINITY else (os.cpu_count() or 1)


def restore_affinity():
    """Undo a thread pin inherited from the environment (OMP_PROC_BIND / GOMP_CPU_AFFINITY set by
    the caller): the engine's host threads, NCCL's and the CUDA driver's inherit the mask of the
    thread that creates them."""
    if _FULL_AFFINITY:
        try:
            os.sched_setaffinity(0, _FULL_AFFINITY)
        except OSError:
            pass


def hbm_peak_gbs():
    """Measured copy bandwidth of this pool's B200s (MEASURED_PEAKS.json), else the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst)"
    except Exception:
        return 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
    `ncu --set full` capture of this round (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the
    .ncu-rep).  None when no capture of the current kernel is committed -- never a constant from an older one."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            d = json.load(fh)
        return float(d["dram_bytes_per_launch"]), d.get("note", "")
    except Exception:
        return None, "no ncu --set full capture of this kernel committed for this round"


def assembly_algorithmic_bytes(a):
    """Extend-add traffic of one factorisation without delays: every child contribution
    entry is read once and its parent entry read and written, (m-n)(m-n+1)/2 entries x 24 B per
    child front, plus the 4 (m-n) index bytes.  Front initialisation (the timed class holds the memset of
    the level's factor storage and k_scatter_a only): 8 B written per entry of L (memset) + 24 B read per
    entry of A (value, int64 source and destination index)."""
    m = (a.rptr[1:] - a.rptr[:-1]).astype(np.float64)
    nc = (a.sptr[1:] - a.sptr[:-1]).astype(np.float64)
    cm = m - nc
    has_parent = np.asarray(a.sparent) <= a.nnodes
    ext = float(((cm * (cm + 1) / 2 * 24 + 4 * cm) * has_parent).sum())
    nz = int(a.ptr[a.n]) - 1
    return {"extend_add": ext, "scatter_a": 24.0 * nz, "zero_L": 8.0 * float((m * nc).sum())}


def hbm_bound_parts(a, inform, solve_ms, tm):
    """Achieved HBM GB/s of the memory-bound parts, algorithmic bytes / measured time: the 1-RHS solve
    (reads L once forward, once backward: 2 x 8 x num_factor bytes) from the device-resident
    solve timing, and assembly / front initialisation from the CUDA-event class times of the profiled
    factorisation (timings()[8..16] = diag, apply, commit, inner, swap, outer, contrib, assemble, init)."""
    peak, src = hbm_peak_gbs()
    out = {"peak": peak, "unit": "GB/s", "peak_source": src}
    t1 = solve_ms.get("1_device")
    if t1:
        b = 2 * 8.0 * inform["num_factor"]
        out["solve_1rhs"] = {"bytes": b, "ms": t1, "achieved": b / (t1 * 1e-3) / 1e9, "frac": b / (t1 * 1e-3) / 1e9 / peak,
                             "note": "fwd + diag_bwd sweeps read L once each (2 x 8 x num_factor bytes)"}
    ab = assembly_algorithmic_bytes(a)
    t_asm, t_init = float(tm[8 + 7]), float(tm[8 + 8])
    if t_asm > 0:
        out["extend_add"] = {"bytes": ab["extend_add"], "ms": t_asm, "achieved": ab["extend_add"] / (t_asm * 1e-3) / 1e9,
                             "frac": ab["extend_add"] / (t_asm * 1e-3) / 1e9 / peak, "kernel": "k_assemble"}
    if t_init > 0:
        bi = ab["scatter_a"] + ab["zero_L"]
        out["init_fronts"] = {"bytes": bi, "ms": t_init, "achieved": bi / (t_init * 1e-3) / 1e9,
                              "frac": bi / (t_init * 1e-3) / 1e9 / peak, "kernel": "cudaMemsetAsync + k_scatter_a"}
    return out


def cpu_reference_arm(args, order):
    """cpu_baseline leg: the reference's own CPU engine (oracle/_ref, unmodified sources) on the SAME workload,
    in a process of its own (`bench.py --impl reference`) so that its OpenMP binding does not touch this
    process; the ordering computed by this process is handed over (METIS once).  Returns the arm's JSON line."""
    import tempfile
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMP_NUM_THREADS", "MASTER_ADDR", "MASTER_PORT")}
    with tempfile.NamedTemporaryFile(suffix=".npy", delete=False) as fh:
        np.save(fh, np.asarray(order, dtype=np.int32))
        opath = fh.name
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                              "--warmup", "0", "--workload", args.workload, "--grid", str(args.grid),
                              "--nrhs", str(args.nrhs), "--order-file", opath], env=env, capture_output=True,
                             text=True, timeout=1500)
    finally:
        os.unlink(opath)
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("reference arm printed no JSON line: " + out.stderr[-300:])


def run_reference(args, rank):
    """The reference's own CPU implementation of the path (src/ssids/cpu, compiled unmodified into oracle/_ref)
    on the box's host cores: ssids_factor of the same workload, then solve_fwd + solve_diag_bwd for 1 and
    nrhs right-hand sides (driver pattern: driver/spral_ssids.F90:419-480)."""
    if rank != 0:
        return
    os.environ.setdefault("OMP_PROC_BIND", "TRUE")     # before libgomp loads (OpenMP tasks bound to cores)
    import oracle_ref
    from spral_b200.ssids import Analysis
    if not oracle_ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libspral_cpu_ref.so was not built"}))
        return
    oracle_ref.ensure_env()
    wl_desc, posdef, scaling_method, (n, ptr, row, val) = make_workload(args)
    order = np.load(args.order_file) if args.order_file else None
    a = Analysis(n, ptr, row, order=order)
    scaling_vec = None
    if scaling_method:
        from spral_b200 import ssids as host
        scaling_vec = host.compute_scaling(a, val, scaling_method)
    cores = host_cores()
    nwarm, nsteps = min(args.warmup, 1), max(1, min(args.steps, REF_MAX_STEPS))
    times, inform, parts, sc = [], None, None, None
    for it in range(nwarm + nsteps):
        if parts is not None:
            for p in parts:
                p.close()
        parts, inform, sc = oracle_ref.ref_factor(a, posdef, val, scaling=scaling_vec, nthreads=cores)
        if it >= nwarm:
            times.append(inform["factor_time"])
    t = float(np.mean(times))
    flops = inform["num_flops"]
    v = flops / t / 1e9
    # solves: the sweeps only (x already permuted).  The reference's solve is a serial loop over the nodes whose only
    # parallelism is the BLAS library's; OpenBLAS threads on these small per-node calls make it 100x SLOWER (measured),
    # so the sweeps run the way the factorisation leaves the library: one BLAS thread.
    blas_threads = int(os.environ.get("SPRAL_B200_REF_SOLVE_BLAS_THREADS", "1"))
    oracle_ref.load().oracle_blas_set_threads(blas_threads)
    solve_s = {}
    rng = np.random.default_rng(0)
    for nr in sorted({1, args.nrhs}):
        x2 = np.asfortranarray(rng.uniform(-1, 1, (n, nr)))
        best = None
        for rep in range(2 if nr == 1 else 1):
            t0 = time.perf_counter()
            for st in parts:
                st.solve("fwd", x2, nr)
            for st in reversed(parts):
                st.solve("diag_bwd", x2, nr)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        solve_s[str(nr)] = best
    for p in parts:
        p.close()
    sample = (f"the whole workload: {nsteps} full factorisation(s) after {nwarm} warm-up (capped at {REF_MAX_STEPS}: one "
              f"takes {t:.0f} s of CPU), same matrix / ordering / options as the GPU arm; solves: fwd + diag_bwd sweeps")
    cpu = {"value": v, "unit": "GFLOP/s", "cores": cores, "kind": "reference", "sample": sample, "seconds": t,
           "solve_seconds": solve_s, "solve_cores": blas_threads,
           "solve_note": "the reference's solve is a serial loop over nodes; BLAS threads = solve_cores",
           "inform": {k: int(inform[k]) for k in ("flag", "num_neg", "num_two", "num_delay", "matrix_rank", "num_factor", "num_flops")}}
    print(json.dumps({
        "impl": "reference", "metric": "ssids_factor FP64 GFLOP/s", "value": v, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": nsteps, "warmup": nwarm, "steps_requested": args.steps,
        "warmup_requested": args.warmup, "ms_per_step": 1e3 * t,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl_desc, "engine": "reference CPU engine src/ssids/cpu (OpenMP tasks + OpenBLAS)",
                   "nparts": int(a.nparts)},
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "solve_ms": {k: 1e3 * x for k, x in solve_s.items()},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=100, help="stencil grid size (BASELINE: 100)")
    ap.add_argument("--workload", default="cfg5", choices=["cfg5", "cfg4", "cfg3", "cfg2"],
                    help="BASELINE config to run (default cfg5 = configs[4], the headline workload)")
    ap.add_argument("--order-file", default=None, help="(reference arm) .npy elimination order computed by the caller")
    ap.add_argument("--nrhs", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    restore_affinity()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    if os.environ.get("SPRAL_B200_ONE_DEVICE"):      # debugging aid: ranks share GPUs (gloo instead of NCCL)
        local_rank = rank % int(os.environ["SPRAL_B200_ONE_DEVICE"])
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("SPRAL_B200_ONE_DEVICE"):
            dist.init_process_group(backend="gloo")
        else:
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    import spral_b200 as sb
    from spral_b200 import _lib, dist as sdist
    lib = _lib.load()

    # ---- set-up (untimed): matrix, ordering, symbolic analysis, subtree partition ----
    wl_desc, posdef, scaling_method, (n, ptr, row, val) = make_workload(args)
    t0 = time.perf_counter()
    ctx = sdist.DistContext(world, rank, local_rank)
    ak = sdist.analyse(ctx, n, ptr, row)
    t_analyse = time.perf_counter() - t0
    a = ak.analysis
    nz = len(val)
    scaling_vec = None
    if scaling_method:                  # host pre-processing of options%scaling, untimed like the analyse phase
        from spral_b200 import ssids as host
        scaling_vec = host.compute_scaling(a, val, scaling_method)
    hval = torch.from_numpy(val).pin_memory()
    dval = hval.cuda(non_blocking=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def factor_once(values_ptr):
        """One ssids_factor over all parts; returns (fkeep, wall seconds, device ms of this rank)."""
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        ev0.record()
        fk = sdist.factor(ctx, ak, posdef, values_ptr, scaling=scaling_vec)   # returns when this rank's streams have drained
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t
        span_ms = ev0.elapsed_time(ev1)                  # device clock, first launch .. last part done
        dev_ms = sum(float(ns.timings()[1]) for ns in fk.numeric if ns is not None)
        barrier()
        return fk, wall, (dev_ms if world == 1 else span_ms)

    # ---- warm-up ----
    fk = None
    for _ in range(args.warmup):
        if fk is not None:
            sdist.free(fk)
        fk, _, _ = factor_once(dval.data_ptr())

    # ---- timed: A resident in HBM ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    walls, launches = [], 0
    for _ in range(args.steps):
        if fk is not None:
            sdist.free(fk)
        fk, wall, dev_ms = factor_once(dval.data_ptr())
        # whole-job time = slowest rank, on the device clock: at N=1 the CUDA events the engine
        # records on its factorisation stream, at N>1 events bracketing this rank's parts
        # (several streams and the waits for other ranks' contribution blocks in between)
        t_step = max_over_ranks(dev_ms / 1e3)
        walls.append(t_step)
        launches += int(sum(float(ns.timings()[6]) for ns in fk.numeric if ns is not None))
    clocks = sampler.stop()
    inform = sdist.reduce_inform(ctx, fk.inform)
    flops = inform["num_flops"]
    t_mean = float(np.mean(walls))
    value = flops / t_mean / 1e9

    # ---- end to end: host (pinned) values through the C ABI, results read back ----
    e2e_t = []
    for _ in range(max(2, min(args.steps, 3))):
        sdist.free(fk)
        fk, wall, _ = factor_once(hval.data_ptr())
        e2e_t.append(max_over_ranks(wall))
    e2e_value = flops / float(np.mean(e2e_t)) / 1e9
    front_bytes = 192 * a.nnodes           # sizeof(Front) (engine.h) per front, read back once per level

    # ---- solves (fkeep%inner_solve): 1 and nrhs right-hand sides ----
    # "device": b and x resident in HBM (CUDA-side permutation, sweeps, un-permutation);
    # "e2e": host arrays in, host arrays out (adds the H2D / D2H copies of b and x).
    solve_ms, bwd_err, bwd_ref1 = {}, None, None
    try:
        from spral_b200 import matrices as M
        A = M.to_scipy(n, ptr, row, val)
        rng = np.random.default_rng(0)
        for nr in (1, args.nrhs):
            X = np.asfortranarray(rng.uniform(-1, 1, (n, nr)))
            X[:, 0] = 1.0
            B = np.asfortranarray(A @ X)
            Bd = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()
            sdist.solve_device(ctx, fk, Bd)                       # warm-up
            barrier()
            t = time.perf_counter()
            Xd = sdist.solve_device(ctx, fk, Bd)
            torch.cuda.synchronize()
            solve_ms[f"{nr}_device"] = 1e3 * max_over_ranks(time.perf_counter() - t)
            barrier()
            t = time.perf_counter()
            Xs = sdist.solve(ctx, fk, B)
            solve_ms[f"{nr}_e2e"] = 1e3 * max_over_ranks(time.perf_counter() - t)
            if nr == 1:
                R = B - A @ Xs
                Xs2 = Xs + sdist.solve(ctx, fk, np.asfortranarray(R))
                if rank == 0:
                    import oracle_ref
                    bwd_err = float(oracle_ref.backward_error(A, Xs, B))
                    bwd_ref1 = float(oracle_ref.backward_error(A, Xs2, B))
    except Exception as e:                                      # solve figures are extra
        solve_ms["error"] = repr(e)

    # ---- roofline of the dominant kernel (profiled extra run, not part of the timing) ----
    roofline = None
    if world == 1:
        lib.spral_ssids_b200_set_profile(1)
        sdist.free(fk)
        fk, _, _ = factor_once(dval.data_ptr())
        lib.spral_ssids_b200_set_profile(0)
        tm = fk.numeric[0].timings()
        peak = float(lib.spral_ssids_b200_fp64_peak_tflops(local_rank))
        ach = float(tm[3]) / (float(tm[2]) * 1e-3) / 1e12 if tm[2] > 0 else 0.0
        traffic, traffic_note = ncu_traffic()
        roofline = {"bound": "tensor",
                    "kernel": "k_update_ws<128,2,4,3,32> UPD_CONTRIB (Schur complement, FP64 DMMA, TMA bulk staged)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak > 0 else None,
                    "traffic": traffic, "traffic_note": traffic_note,
                    "launches": int(tm[4]), "kernel_ms_per_step": float(tm[2]),
                    "kernel_flops_per_step": float(tm[3]),
                    "peak_source": "FP64 DMMA issue loop measured live on this GPU "
                                   "(MEASURED_PEAKS.json has no FP64 figure; cuBLAS DGEMM 8192^3 measured 35.9 TF/s on this pool)"}
        # the 64-RHS solve against the same tensor roofline (4 x num_factor x nrhs flops)
        tn = solve_ms.get(f"{args.nrhs}_device")
        if tn and args.nrhs > 1:
            fl = 4.0 * inform["num_factor"] * args.nrhs
            roofline["solve_nrhs"] = {"nrhs": args.nrhs, "flops": fl, "ms": tn, "achieved": fl / (tn * 1e-3) / 1e12,
                                      "frac": fl / (tn * 1e-3) / 1e12 / peak if peak > 0 else None, "unit": "TFLOP/s"}

    # ---- the HBM-bound parts (north_star: achieved GB/s of assembly and solves against the copy bandwidth) ----
    hbm_rooflines = None
    if world == 1 and roofline is not None:
        try:
            hbm_rooflines = hbm_bound_parts(a, inform, solve_ms, tm)
        except Exception as e:                                  # extra figures only
            hbm_rooflines = {"error": repr(e)}

    out = {
        "metric": "ssids_factor FP64 GFLOP/s", "value": value, "unit": "GFLOP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_mean,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl_desc,
                   "parallelism": f"subtree-partition x{world}" if world > 1 else "1 GPU",
                   "nparts": int(a.nparts), "l2": "inputs larger than L2 (factor storage %.1f GB)" % (inform["num_factor"] * 8 / 1e9),
                   "timer": "CUDA events: on the factorisation stream (N=1); bracketing each rank's parts, max over ranks (N>1)"},
        "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": int(nz * 8),
                "d2h_bytes_per_step": int(front_bytes)},
        "gpu_launches": launches,
        "clocks": clocks,
        "inform": {k: int(inform[k]) for k in ("flag", "num_neg", "num_two", "num_delay", "matrix_rank", "num_factor", "num_flops", "maxfront")},
        "solve_ms": {str(k): v for k, v in solve_ms.items()},
        "backward_error": bwd_err, "backward_error_after_1_refinement": bwd_ref1,
        "analyse_s": t_analyse,
    }
    if roofline:
        out["roofline"] = roofline
    if hbm_rooflines:
        out["roofline_hbm_parts"] = hbm_rooflines

    # ---- CPU baseline: the reference engine on a bounded sample (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import oracle_ref
            if oracle_ref.available():
                ref = cpu_reference_arm(args, a.order)
                if "cpu_baseline" not in ref:
                    raise RuntimeError(ref.get("unavailable", "no cpu_baseline in the reference arm's line"))
                out["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}
    sdist.free(fk)
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
