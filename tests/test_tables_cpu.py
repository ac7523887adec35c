"""Host-side exhaustive checks of the constant tables the CUDA kernels use (no GPU needed):
the cell LUT reproduces the double-precision curve for every possible matrix output, and the
encode path's three-instruction constant division equals IEEE division for every 8-bit RGB."""
from metalbt709decoder_b200 import _lib


def test_tables_selftest():
    assert _lib.lib().bt709_selftest_host() == 0
