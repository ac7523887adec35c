"""examples/roundtrip.c: the reference-facing host API used from plain C99."""
import os
import subprocess

import pytest

from metalbt709decoder_b200 import _lib

EXE = os.path.join(os.path.dirname(_lib.LIB_PATH), "roundtrip_example")


def test_example_builds_and_has_no_cpu_fallback():
    """built by `make all` with gcc -std=c99 -pedantic against include/metal_bt709_decoder.h; without a
    CUDA device it must refuse to run (exit 3), never compute on the CPU"""
    _lib.lib()
    assert os.path.exists(EXE), "make -C metalbt709decoder_b200/csrc all builds the example"
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_example_roundtrip")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["apple", "srgb"])
def test_example_roundtrip(mode):
    """encode -> decode round trip through the C API: gray and colour ramps come back within the
    reference's own round-trip error (CoreImageMetalFilterTests.m:541-551: |delta| <= 5)"""
    r = subprocess.run([EXE, "512", "256", mode], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "kernels launched: 2" in r.stdout, r.stdout
