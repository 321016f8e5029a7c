"""Host-side environment checks (CPU only).

Regression test for a scheduling bug of the first multi-GPU runs: bench.py exported
OMP_PROC_BIND=TRUE for the reference CPU engine, libgomp (loaded by torch) then pinned
the initial thread of every rank to the first CPU of the mask, and every host thread of
every rank inherited it -- N ranks shared one core (analyse 15 s -> 141 s at N = 8)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.skipif(not hasattr(os, "sched_getaffinity"), reason="needs sched_getaffinity")


def _run(code, **env):
    e = {k: v for k, v in os.environ.items() if not k.startswith("OMP_")}
    e.update(env)
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()[-1]


def test_bench_b200_arm_does_not_pin_host_threads():
    code = ("import os; full = os.sched_getaffinity(0)\n"
            "import bench, torch, spral_b200\n"
            "import threading; seen = []\n"
            "t = threading.Thread(target=lambda: seen.append(os.sched_getaffinity(0))); t.start(); t.join()\n"
            "print(int(os.sched_getaffinity(0) == full and seen[0] == full))")
    assert _run(code) == "1"


def test_bench_restores_an_inherited_pin():
    if len(os.sched_getaffinity(0)) < 2:
        pytest.skip("one CPU only")
    code = ("import os; full = os.sched_getaffinity(0)\n"
            "import bench, torch\n"              # the caller's OMP_PROC_BIND pins the main thread here
            "pinned = os.sched_getaffinity(0)\n"
            "bench.restore_affinity()\n"
            "print(int(os.sched_getaffinity(0) == full), int(len(pinned) < len(full)))")
    ok, was_pinned = _run(code, OMP_PROC_BIND="TRUE").split()
    assert ok == "1"


def test_cpu_baseline_leg_runs_in_its_own_process():
    """The cpu_baseline leg is the reference arm in a subprocess; its line carries the keys
    bench.py copies."""
    import json
    import oracle_ref
    if not oracle_ref.available():
        pytest.skip("oracle/_ref not built")
    code = ("import bench, json, argparse, numpy as np\n"
            "from spral_b200 import matrices as M\n"
            "from spral_b200.ssids import Analysis\n"
            "args = argparse.Namespace(workload='cfg5', grid=12, nrhs=4)\n"
            "n, ptr, row, val = M.stencil_3d_27pt(12, shift=13.0)\n"
            "a = Analysis(n, ptr, row)\n"
            "print(json.dumps(bench.cpu_reference_arm(args, a.order)))")
    ref = json.loads(_run(code).strip().splitlines()[-1])
    cb = ref["cpu_baseline"]
    assert ref["impl"] == "reference" and cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0
    assert ref["e2e"]["h2d_bytes_per_step"] == 0
    assert "12^3" in ref["config"]["workload"]                       # the SAME workload as the GPU arm (same_config)
    assert set(cb["solve_seconds"]) == {"1", "4"} and cb["inform"]["matrix_rank"] == 1728


def test_trivial_matrix_n0():
    """ssids_analyse / factor / solve on the empty matrix return immediately with rank 0
    (src/ssids/ssids.f90:212-218, 848-852, 1193); no GPU is touched."""
    import numpy as np
    import spral_b200 as sb
    ak = sb.analyse(0, np.array([1], dtype=np.int64), np.zeros(0, dtype=np.int32))
    assert ak.analysis.nnodes == 0 and ak.analysis.nparts == 0 and ak.subtrees == []
    fk = sb.factor(ak, False, np.zeros(0))
    assert fk.inform["flag"] == 0 and fk.inform["matrix_rank"] == 0 and fk.numeric == []
    assert sb.solve(fk, np.zeros(0)).shape == (0,)
