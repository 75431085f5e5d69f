"""Error convention of the C ABI (return code for API / CUDA errors, algorithmic outcome only in the
per-instance status): exercised here on the emulator build of the same ipddp_api.cu; tests/test_gpu_parity.py runs the same checks on
the GPU through the product library."""
import os
import sys

import helpers

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))


def test_error_convention_emulated():
    import build_emu
    from ipddp_b200 import _lib
    helpers.api_error_convention(_lib.Lib(build_emu.build()))
