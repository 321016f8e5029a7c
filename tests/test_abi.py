"""CPU-only checks of the drop-in boundary: the shared library loads, exports
every symbol include/spral_ssids_b200.h declares, and the ctypes images of the
interoperable structs have the C layout."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT
from spral_b200 import _lib

HEADER = os.path.join(ROOT, "include", "spral_ssids_b200.h")


COMPAT = os.path.join(ROOT, "include", "spral_ssids_compat.h")


def declared_functions(header=HEADER):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spral_ssids_\w+)\s*\(", src)))


def test_header_declares_what_python_binds():
    assert declared_functions() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, missing


def test_load_binds_all():
    lib = _lib.load()
    for name in _lib.SYMBOLS:
        assert getattr(lib, name) is not None


def test_struct_layouts_match_c(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "spral_ssids_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(struct spral_ssids_b200_options), sizeof(struct spral_ssids_b200_stats),'
                    'sizeof(struct spral_ssids_b200_contrib), offsetof(struct spral_ssids_b200_options,small_subtree_threshold),'
                    'offsetof(struct spral_ssids_b200_stats,num_flops), offsetof(struct spral_ssids_b200_contrib,owner_ptr));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).split()
    got = [int(v) for v in out]
    want = [C.sizeof(_lib.Options), C.sizeof(_lib.Stats), C.sizeof(_lib.Contrib),
            _lib.Options.small_subtree_threshold.offset, _lib.Stats.num_flops.offset,
            _lib.Contrib.owner_ptr.offset]
    assert got == want


def test_options_layout_matches_reference_cpu_factor_options():
    # cpu_factor_options (src/ssids/cpu/cpu_iface.hxx:24-34): int, bool, double, double,
    # double, int64, int, int(enum), int(enum)
    o = _lib.Options
    assert [f[0] for f in o._fields_] == ["print_level", "action", "small", "u", "multiplier",
                                          "small_subtree_threshold", "cpu_block_size",
                                          "pivot_method", "failed_pivot_method"]
    assert C.sizeof(o) == 56


def test_compat_header_symbols_exported_and_layout_matches_reference_sizes():
    """include/spral_ssids_compat.h (the spral_ssids.h subset): every function is
    exported; the two interoperable structs have the sizes of the reference's
    (include/spral_ssids.h:15-57: options 176 bytes, inform 152 bytes)."""
    lib = C.CDLL(_lib.LIB_PATH)
    names = [f for f in declared_functions(COMPAT) if f.startswith("spral_ssids_")]
    assert len(names) == 15        # every function of spral_ssids.h:70-127
    assert not [f for f in names if not hasattr(lib, f)]
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "spral_ssids_compat.h"\nint main(void){printf("%zu %zu\\n",'
                             'sizeof(struct spral_ssids_options), sizeof(struct spral_ssids_inform));return 0;}\n')
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        a, b = map(int, subprocess.check_output([exe]).split())
    ref = os.path.join("/root/reference/include/spral_ssids.h")
    if os.path.exists(ref):
        with tempfile.TemporaryDirectory() as td:
            src = os.path.join(td, "r.c")
            open(src, "w").write('#include <stdio.h>\n#include "spral_ssids.h"\nint main(void){printf("%zu %zu\\n",'
                                 'sizeof(struct spral_ssids_options), sizeof(struct spral_ssids_inform));return 0;}\n')
            exe = os.path.join(td, "r")
            subprocess.check_call(["gcc", "-I", "/root/reference/include", src, "-o", exe])
            ra, rb = map(int, subprocess.check_output([exe]).split())
        assert (a, b) == (ra, rb)
    assert (a, b) == (176, 152)
