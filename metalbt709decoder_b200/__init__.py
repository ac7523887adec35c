"""metalbt709decoder_b200 — B200-native BT.709 4:2:0 <-> sRGB BGRA8 pixel path.

The product is libbt709_b200.so (CUDA kernels for sm_100a behind the C ABI of
include/bt709_b200.h, plus the reference-facing host layer of include/metal_bt709_decoder.h).
This package binds it with ctypes; it contains no arithmetic and no CPU fallback.
"""
from . import _lib  # noqa: F401
from .device_api import (BT709Context, BT709Error, GAMMA_APPLE, GAMMA_BT709, GAMMA_LINEAR, GAMMA_NAMES,  # noqa: F401
                         GAMMA_SRGB)
from .sharding import frame_slice  # noqa: F401

__all__ = ["BT709Context", "BT709Error", "GAMMA_APPLE", "GAMMA_SRGB", "GAMMA_LINEAR", "GAMMA_BT709", "GAMMA_NAMES",
           "frame_slice"]
