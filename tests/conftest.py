import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
# the reference CPU engine needs these before libgomp starts (src/ssids/ssids.f90:1448-1452)
os.environ.setdefault("OMP_CANCELLATION", "TRUE")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


# TEST INFRASTRUCTURE: with SPRAL_B200_EMU_LIB=<build/emu/libspral_ssids_b200_emu.so> (tests/emu/build_emu.py) the tests
# talk to the engine's real sources compiled for the CPU on the SIMT emulator of tests/emu -- a way to run the logic of
# the GPU tests in a GPU-less container.  Only the tests do this; the package itself knows one library, the CUDA one.
if os.environ.get("SPRAL_B200_EMU_LIB"):
    from spral_b200 import _lib as _emu_lib
    _emu_lib.LIB_PATH = os.environ["SPRAL_B200_EMU_LIB"]
