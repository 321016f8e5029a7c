"""CPU checks of device-code bodies that are written once and compiled twice.

spral_b200/csrc/diag_warp.cuh (the one-warp 32 x 32 LDL^T / Cholesky of k_diag_w and of the chain
kernel) runs on the fiber emulator of tests/emu; solve_wide.h on host threads (tests/c/emu.h); the
pivoting state machine, the two-stream look-ahead and the distributed-front protocol are model-checked."""
import os
import subprocess

from conftest import ROOT

REF = "/root/reference"


def test_wide_solve_bodies_match_plain_sweeps():
    """spral_b200/csrc/solve_wide.h (256-column forward / backward sweeps, T and G kernels)
    on host threads against gather / substitute / scatter over one front."""
    out = os.path.join(ROOT, "build", "tests")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "solve_wide_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe,
                           os.path.join(ROOT, "tests", "c", "solve_wide_emu.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "solve_wide_emu: 0 failures" in r.stdout


def test_one_warp_diag_block_factorisation():
    """spral_b200/csrc/diag_warp.cuh (body of k_diag_w and of the chain kernel's 32 x 32 steps) on the fiber emulator
    of tests/emu: P A P^T = L D L^T, mirrored L*D, ties, zero pivots, short blocks, Cholesky, not positive definite."""
    import pytest
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    out = os.path.join(ROOT, "build", "tests")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "diag_warp_emu")
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-I" + emu, "-I" + os.path.join(ROOT, "spral_b200", "csrc"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + cuda_inc, "-include", os.path.join(emu, "cuda_emu.h"),
                           "-o", exe, os.path.join(ROOT, "tests", "c", "diag_warp_emu.cpp"),
                           os.path.join(emu, "cuda_emu_rt.cpp"), "-lrt"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert ", 0 failures" in r.stdout


def test_pivoting_protocol_model_check():
    """tests/c/pivot_state_emu.cpp: the device state machine (spral_b200/csrc/pivot_state.h, the code the
    kernels run) against the host mirror of factor_fronts, over thousands of random levels with failed
    block columns, unsplittable 2x2 pivots, passes, delays and accepted / given-up / rolled-back
    speculative segments: no divergence, every panel complete, termination, consistent statistics."""
    import pytest
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    out = os.path.join(ROOT, "build", "tests")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "pivot_state_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c", "pivot_state_emu.cpp")])
    r = subprocess.run([exe, "6000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout


def test_two_stream_lookahead_schedule_model_check():
    """tests/c/lookahead_race_emu.cpp: factor_fronts' two-stream schedule (panel kernels on the main stream, look-ahead
    bulk updates on the second, ordered only by the per-panel host sync and the ev_bulk / ev_bulk_all waits) restated
    launch for launch on the shared state machine, every launch with the footprint of the real kernel: no unordered
    pair of launches conflicts, every column has every update when its block is factorised -- with and without the
    speculative segments, with failed block columns, passes and delays; four injected faults must be detected."""
    import pytest
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    out = os.path.join(ROOT, "build", "tests")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "lookahead_race_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c", "lookahead_race_emu.cpp")])
    r = subprocess.run([exe, "4000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout and "detected in 0 " not in r.stdout and "/ 0 (" not in r.stdout, r.stdout


def test_distributed_top_front_protocol_model_check():
    """tests/c/dist_front_emu.cpp: the protocol PLANNED for splitting a top front between an owner and a helper GPU
    (not built yet): flags, panel buffers with back-pressure, blocks coming back just in time, draining
    on the first failed pivot.  No deadlock, no unordered conflicting pair over the three streams, every column fully
    updated; four injected protocol faults must be detected."""
    import pytest
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    out = os.path.join(ROOT, "build", "tests")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "dist_front_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c", "dist_front_emu.cpp")])
    r = subprocess.run([exe, "1500"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout and "detected in 0 " not in r.stdout and "/ 0 (" not in r.stdout, r.stdout


def test_build_without_split_front_has_no_trace_of_it():
    """spral_b200/csrc/split_front.h (distributed top front) is compiled in by default and switched on at run time;
    `make NOSPLIT=1` must give a library without a token of it (the preprocessed subtree.cu is checked)."""
    csrc = os.path.join(ROOT, "spral_b200", "csrc")
    pre = subprocess.run(["g++", "-x", "c++", "-E", "-P", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                          "-I/usr/local/cuda/include", os.path.join(csrc, "subtree.cu")], capture_output=True, text=True)
    assert pre.returncode == 0, pre.stderr[-2000:]
    for token in ("SplitOwner", "SplitShm", "split_now", "split_helper_serve", "shm_open"):
        assert token not in pre.stdout, token
    with open(os.path.join(csrc, "Makefile")) as fh:
        assert "-DSPRAL_B200_SPLIT" in fh.read()


def test_split_front_host_code_on_a_cuda_mock():
    """tests/c/split_front_emu.cpp: the real spral_b200/csrc/split_front.h (SplitOwner, split_helper_serve, the
    shared-memory protocol) with the owner and the helper as two threads over a mock of the CUDA runtime calls it makes
    (in-order worker-thread streams, host-callback flags, pointer-carrying IPC handles, UPD_EXPLICIT as a triple loop with
    the kernel's region / tile semantics): ten fronts -- square, with contribution rows, Cholesky flavour, odd sizes,
    "failed pivot" (drain) at the first / middle / last panel, too small to split -- reproduce the unsplit run bit for bit."""
    out = os.path.join(ROOT, "build", "tests")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "split_front_emu")
    cuda_inc = "/usr/local/cuda/include"
    import pytest
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-DSPRAL_B200_SPLIT", "-I" + cuda_inc,
                           "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c", "split_front_emu.cpp"), "-lrt"])
    # the mock's owner loop hands every far block to the helper (the owner's own share of the blocks is covered by
    # the end-to-end emulation, tests/emu/split_check.py)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=dict(os.environ, SPRAL_B200_SPLIT_OWNER_SHARE="0"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "split_front_emu: 10 cases, 0 failures" in r.stdout
