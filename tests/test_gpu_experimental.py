"""Opt-in kernel variants against the default engine (GPU; runs only with
SPRAL_B200_EXPERIMENTAL_TESTS=1 because the variants have not been on a B200 yet).

Each variant is selected by an environment variable that the library reads once per
process, so both sides run tools/dump_factor.py in a process of their own.  A variant
that only changes HOW the same arithmetic is scheduled must reproduce pivot order,
D^-1, inform and the solution bit for bit:
  SPRAL_B200_DIAG_V2=1     4-warp diagonal-block kernel (diag_block.h; the body is
                           checked on the CPU by tests/test_kernel_emulation.py)
  SPRAL_B200_BULK_PRIO=1   look-ahead bulk update one tile per CTA on a low-priority
                           stream instead of a capped persistent grid
  SPRAL_B200_CTILE_BLOCK=4 Schur-complement tiles in blocked order (L2 reuse)
one that factorises the panels speculatively in 128-column segments (panel_v2.h; other
but equally valid pivots where entries tie, sums in another order):
  SPRAL_B200_PANEL_V2=1    inertia, rank and flops identical; delays close; solutions to rounding
and one that changes the order of the sums of the solves (same factors; solutions
agree to rounding):
  SPRAL_B200_SOLVE_WIDE=1  256-column sweeps on the levels of large fronts (solve_wide.h;
                           bodies checked on the CPU by tests/test_kernel_emulation.py)
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SPRAL_B200_EXPERIMENTAL_TESTS") != "1",
                                 reason="set SPRAL_B200_EXPERIMENTAL_TESTS=1 to run the opt-in kernel variants")]


def _dump(tmp_path, tag, **env):
    out = str(tmp_path / f"{tag}.npz")
    e = dict(os.environ)
    for k in ("SPRAL_B200_DIAG_V2", "SPRAL_B200_BULK_PRIO", "SPRAL_B200_CTILE_BLOCK", "SPRAL_B200_SOLVE_WIDE",
              "SPRAL_B200_SOLVE_WIDE_MIN", "SPRAL_B200_PANEL_V2"):
        e.pop(k, None)
    e.update(env)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "dump_factor.py"), out], env=e, timeout=900)
    return np.load(out)


@pytest.fixture(scope="module")
def baseline(tmp_path_factory):
    return _dump(tmp_path_factory.mktemp("base"), "base")


@pytest.mark.parametrize("var,value", [("SPRAL_B200_DIAG_V2", "1"), ("SPRAL_B200_BULK_PRIO", "1"),
                                       ("SPRAL_B200_CTILE_BLOCK", "4")])
def test_variant_reproduces_default_engine_bit_for_bit(tmp_path, baseline, var, value):
    got = _dump(tmp_path, var, **{var: value})
    assert sorted(got.files) == sorted(baseline.files)
    for k in baseline.files:
        assert np.array_equal(baseline[k], got[k], equal_nan=True), (var, k)


@pytest.mark.parametrize("wide_min", ["2", "8"])
def test_wide_solve_agrees_with_narrow_sweeps(tmp_path, baseline, wide_min):
    got = _dump(tmp_path, "wide" + wide_min, SPRAL_B200_SOLVE_WIDE="1", SPRAL_B200_SOLVE_WIDE_MIN=wide_min)
    for k in baseline.files:
        if k.endswith("/x") or k.endswith("/x5"):
            scale = np.abs(baseline[k]).max()
            assert np.abs(baseline[k] - got[k]).max() <= 1e-9 * scale, k
        else:                                            # the factorisation is untouched
            assert np.array_equal(baseline[k], got[k], equal_nan=True), k


def test_speculative_panel_segments_agree_with_step_by_step_path(tmp_path, baseline):
    got = _dump(tmp_path, "panel_v2", SPRAL_B200_PANEL_V2="1")
    for k in baseline.files:
        if k.endswith("/inform"):
            b, g = baseline[k], got[k]      # flag, num_delay, num_neg, num_two, matrix_rank, num_factor, num_flops
            assert g[0] == b[0] and g[2] == b[2] and g[4] == b[4], (k, b, g)
            assert abs(int(g[1]) - int(b[1])) <= 8 + 0.25 * int(b[1]), (k, b, g)
            if b[1] == 0 and g[1] == 0:
                assert g[5] == b[5] and g[6] == b[6], (k, b, g)
        elif k.endswith("/x") or k.endswith("/x5"):
            scale = np.abs(baseline[k]).max()
            assert np.abs(baseline[k] - got[k]).max() <= 1e-7 * scale, k
