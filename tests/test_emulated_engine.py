"""The engine's REAL sources on a CPU: tests/emu compiles subtree.cu, factor_kernels.cu and solve_kernels.cu with g++ on
top of a small SIMT emulator (every CUDA thread of a CTA on a fiber, __syncthreads / shuffles as rendezvous points,
__shared__ as static storage, the CUDA runtime answered synchronously; only gemm_dmma.cu -- inline PTX -- is replaced
by loops over the same regions and tiles).  The library is loaded BY TESTS ONLY (tests/conftest.py,
SPRAL_B200_EMU_LIB); the package knows the CUDA library and nothing else.

This runs the logic of the GPU test files in a GPU-less container: host scheduling, assembly, pivoting kernels, delays,
solves -- and the alternative code paths -- against the oracle.  It is not a parity claim (parity is measured on the
B200, `-m gpu`); it is how changes are checked before GPU minutes are spent on them.  The CPU suite runs a slice of a
few minutes."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

EMU_LIB = os.path.join(ROOT, "build", "emu", "libspral_ssids_b200_emu.so")


@pytest.fixture(scope="module")
def emu_env():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "emu", "build_emu.py")], stdout=subprocess.DEVNULL)
    env = dict(os.environ)
    env.update(SPRAL_B200_EMU_LIB=EMU_LIB, OMP_CANCELLATION="TRUE")
    return env


def _pytest(env, args, timeout):
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + args, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return r.stdout


def test_slice_of_the_gpu_parity_suite_on_the_emulator(emu_env):
    out = _pytest(emu_env, ["tests/test_gpu_parity.py", "-k",
                            "(dense_fronts and (31 or 33 or 100)) or (doctored and 64) or singular_matrix or not_positive_definite "
                            "or solve_jobs or enquire or multi_part or index_maps"], 900)
    assert " passed" in out and "failed" not in out, out[-500:]


def test_alternative_code_paths_against_the_defaults_on_the_emulator(emu_env):
    """tests/test_gpu_paths.py (step-by-step panels, 32-column solve kernels, scheduling switches, sweeps without
    look-ahead against the defaults) on a small dense front."""
    env = dict(emu_env, SPRAL_B200_DUMP_CASES="dense_391_indef")
    out = _pytest(env, ["tests/test_gpu_paths.py"], 2400)
    assert "4 passed" in out, out[-500:]


def test_solve_paths_of_many_right_hand_sides_and_large_fronts_on_the_emulator(emu_env):
    """tests/emu/solve_check.py: tensor-core G kernels, look-ahead inside a sweep, inverse diagonal blocks (L D L^T and
    Cholesky), half-height T kernels, chunks of 64 + 16 + 4 right-hand sides (the switched-off variants are compared
    by tests/test_gpu_paths.py, on the emulator above and on the GPU)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "solve_check.py")], env=emu_env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.count(" OK") == 5, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.fixture(scope="module")
def emu_split_lib(emu_env):
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "emu", "build_emu.py"), "split"], stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "build", "emu_split", "libspral_ssids_b200_emu_split.so")


@pytest.mark.parametrize("n,kind,helpers,expect", [
    (1300, "indef", 1, "drain at panel"),
    pytest.param(1300, "posdef", 2, "front closed",           # two helpers + the owner's share
                 marks=pytest.mark.skipif(os.environ.get("SPRAL_B200_SLOW_TESTS") != "1", reason="2 minutes: set SPRAL_B200_SLOW_TESTS=1"))])
def test_distributed_top_front_end_to_end_on_the_emulator(emu_env, emu_split_lib, n, kind, helpers, expect):
    """csrc/split_front.h and its hooks in factor_fronts with the owner and the helper
    as two threads of one process on the emulator: a front that is split to its end (Cholesky) and one whose split is
    drained by a failed pivot give the factors and solutions of the unsplit run bit for bit."""
    env = dict(emu_env, SPRAL_B200_EMU_LIB=emu_split_lib, SPRAL_B200_TRACE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "split_check.py"), str(n), kind, str(helpers)], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "BITWISE IDENTICAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert expect in r.stderr, r.stderr[-2000:]


@pytest.mark.skipif(os.environ.get("SPRAL_B200_SLOW_TESTS") != "1", reason="3 minutes: set SPRAL_B200_SLOW_TESTS=1")
@pytest.mark.parametrize("grid,kind,ranks", [(28, "stencil", 2), (34, "lap", 3)])
def test_two_process_gpu_path_with_split_on_the_emulator(emu_env, emu_split_lib, grid, kind, ranks):
    """spral_b200/dist.py on its real GPU code path (GpuEngine, contribution blocks by "CUDA IPC", the distributed top
    front between an owner and a helper rank) as two processes over gloo, "device" memory in POSIX shared memory
    (tests/emu/dist_split_check.py): inertia and statistics equal the single-process run, solutions agree; the indefinite
    problem drains the split at a failed pivot, the Cholesky problem runs it to its end (4 panels out, 4 blocks back)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "dist_split_check.py"), str(grid), kind, str(ranks)],
                       env=dict(os.environ), capture_output=True, text=True, timeout=1700)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("client,ok", [("ssids_capi_check.c", "CAPI OK"), ("ssids_capi_coord_check.c", "CAPI COORD OK")])
def test_c_clients_of_the_spral_ssids_h_interface_on_the_emulator(emu_env, tmp_path, client, ok):
    """The C clients of include/spral_ssids_compat.h (the reference's C interface: analyse / analyse_coord, factor,
    factor_ptr32, solve1, enquire, orderings, scalings) linked against the emulated library."""
    exe = tmp_path / "client"
    libdir = os.path.dirname(EMU_LIB)
    subprocess.check_call(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", client), "-o", str(exe),
                           "-L", libdir, "-lspral_ssids_b200_emu", "-lm", f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([str(exe)], env=emu_env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and ok in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
