"""CPU-side checks of the boundary: the shared library loads, exports every symbol the public
headers declare, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import metalbt709decoder_b200 as pkg
from metalbt709decoder_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.findall(r"BT709_API\s+[\w\s\*]+?\b(\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.LIB_PATH) if os.path.exists(_lib.LIB_PATH) else _lib.lib()
    names = _declared("bt709_b200.h") + _declared("metal_bt709_decoder.h") + _declared("y4m_b200.h")
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported"
    # and the ctypes table covers exactly the declared set
    assert set(names) == set(_lib.PROTOTYPES)


def test_status_strings():
    L = _lib.lib()
    for s in range(0, -9, -1):
        assert L.bt709_status_string(s) and b"unknown" not in L.bt709_status_string(s)
    assert b"unknown" in L.bt709_status_string(-99)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(pkg.BT709Error) as e:
        pkg.BT709Context(0)
    assert e.value.status == -6  # BT709_E_CUDA


def test_null_arguments_are_rejected_not_crashed():
    L = _lib.lib()
    assert L.bt709_ctx_create(0, None) == -1
    assert L.bt709_decode_nv12(None, 0, None, None, None, None, 2, 2, 1, None, 0) == -1
    assert L.bt709_rescale_bgra(None, None, 2, 2, None, 2, 2, 1, None, 0) == -1
    assert L.bt709_encode_bgra_to_nv12(None, 1, 0, None, None, None, 2, 2, 1, None, 0) == -1
    assert L.bt709_convert_packed444(None, None, None, 2, 2, None, 0) == -1
    assert L.bt709_convert_packed444_gamma(None, 3, None, None, 2, 2, None, 0) == -1
    assert L.bt709_unconvert_packed444_gamma(None, 3, None, None, 2, 2, None, 0) == -1
    assert L.bt709_decode_rescale_nv12_fused(None, 0, None, None, None, None, 2, 2, 4, 4, 1, None, 0) == -1
    assert L.bt709_launch_count(None) == 0


def test_host_layer_guards_without_device():
    """BOOL-style error behaviour that needs no GPU (BGRAToBT709Converter.m:41-46, 52-54)."""
    import numpy as np
    from metalbt709decoder_b200 import renderer as R
    mrc = R.MetalRenderContext(0)
    px = np.zeros(6, np.uint32)
    ok, _ = R.BGRAToBT709Converter.convert(mrc, px, 3, 2, R.BGRAToBT709ConverterSoftware)
    assert ok is False  # odd width
    ok, _ = R.BGRAToBT709Converter.convert(mrc, np.zeros(4, np.uint32), 2, 2, R.BGRAToBT709ConverterMetal)
    assert ok is False  # Metal back-end is FALSE in the reference as well
    ok, _ = R.BGRAToBT709Converter.unconvert(mrc, np.zeros(4, np.uint32), 2, 2, R.BGRAToBT709ConverterVImage)
    assert ok is False
    pb = R.BGRAToBT709Converter.createCoreVideoYCbCrBuffer(4, 2)
    assert pb.y.shape == (2, 4) and pb.cbcr.shape == (1, 4)
    assert R.BGRAToBT709Converter.setBT709Attributes(pb)
    assert pb.struct.yCbCrMatrix == R.kMBDYCbCrMatrix_ITU_R_709_2


def test_frame_sharding():
    from metalbt709decoder_b200.sharding import frame_slice
    for F in (1, 7, 600, 2400):
        for N in (1, 2, 4, 8):
            spans = [frame_slice(F, r, N) for r in range(N)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-F // N)
    with pytest.raises(ValueError):
        frame_slice(10, 3, 2)
