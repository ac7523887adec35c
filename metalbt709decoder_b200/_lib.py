"""ctypes binding of libbt709_b200.so (include/bt709_b200.h + include/metal_bt709_decoder.h).

The shared library is the product; this module only loads it and declares prototypes.
There is no fallback of any kind: if the library is missing or no CUDA device is present
the import (respectively context creation) fails loudly.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BT709_B200_LIB") or os.path.join(_HERE, "libbt709_b200.so")  # env: kernel experiments only
CSRC = os.path.join(_HERE, "csrc")

c_void_p, c_int, c_size_t, c_u32p = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)


class Plane(ctypes.Structure):
    """bt709_plane"""
    _fields_ = [("ptr", c_void_p), ("pitch", c_size_t), ("frame_stride", c_size_t)]


class MBDPixelBuffer(ctypes.Structure):
    _fields_ = [("width", c_int), ("height", c_int), ("y", c_void_p), ("yBytesPerRow", c_size_t),
                ("cbcr", c_void_p), ("cbcrBytesPerRow", c_size_t), ("yCbCrMatrix", c_int),
                ("transferFunction", c_int), ("ownsMemory", c_int)]


class MBDTexture(ctypes.Structure):
    _fields_ = [("width", c_int), ("height", c_int), ("pixels", c_void_p), ("bytesPerRow", c_size_t),
                ("isSRGB", c_int), ("ownsMemory", c_int)]


def build(verbose=False):
    """Compile the library in-tree with nvcc for sm_100a (no GPU needed)."""
    r = subprocess.run(["make", "-C", CSRC, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libbt709_b200.so failed")


def _needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    srcs += [os.path.join(_HERE, "..", "include", f) for f in ("bt709_b200.h", "metal_bt709_decoder.h", "y4m_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


_lib = None

PP = ctypes.POINTER(Plane)
PPB = ctypes.POINTER(MBDPixelBuffer)
PTX = ctypes.POINTER(MBDTexture)

# name -> (restype, argtypes); every symbol the two public headers declare
PROTOTYPES = {
    # include/bt709_b200.h
    "bt709_ctx_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "bt709_ctx_destroy": (None, [c_void_p]),
    "bt709_ctx_device": (c_int, [c_void_p]),
    "bt709_last_cuda_error": (c_int, [c_void_p]),
    "bt709_status_string": (ctypes.c_char_p, [c_int]),
    "bt709_launch_count": (ctypes.c_ulonglong, [c_void_p]),
    "bt709_selftest_host": (c_int, []),
    "bt709_decode_nv12": (c_int, [c_void_p, c_int, PP, PP, PP, PP, c_int, c_int, c_int, c_void_p, c_int]),
    "bt709_rescale_bgra": (c_int, [c_void_p, PP, c_int, c_int, PP, c_int, c_int, c_int, c_void_p, c_int]),
    "bt709_decode_rescale_nv12": (c_int, [c_void_p, c_int, PP, PP, PP, PP, PP, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p, c_int]),
    "bt709_decode_rescale_nv12_fused": (c_int, [c_void_p, c_int, PP, PP, PP, PP, c_int, c_int, c_int, c_int, c_int,
                                                c_void_p, c_int]),
    "bt709_encode_bgra_to_nv12": (c_int, [c_void_p, c_int, c_int, PP, PP, PP, c_int, c_int, c_int, c_void_p, c_int]),
    "bt709_convert_packed444": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int]),
    "bt709_unconvert_packed444": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int]),
    "bt709_convert_packed444_gamma": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int]),
    "bt709_unconvert_packed444_gamma": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int]),
    "bt709_copy_packed444_to_nv12": (c_int, [c_void_p, c_void_p, PP, PP, c_int, c_int, c_void_p, c_int]),
    # include/y4m_b200.h
    "y4m_open_file": (c_void_p, [ctypes.c_char_p]),
    "y4m_write_header": (c_int, [c_void_p, c_void_p]),
    "y4m_write_frame": (c_int, [c_void_p, c_void_p]),
    "y4m_read_header": (c_int, [c_void_p, c_void_p]),
    "y4m_read_frame": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bt709_cbcr_to_planar": (c_int, [c_void_p, PP, PP, PP, c_int, c_int, c_int, c_void_p, c_int]),
    "bt709_planar_to_cbcr": (c_int, [c_void_p, PP, PP, PP, c_int, c_int, c_int, c_void_p, c_int]),
    # include/metal_bt709_decoder.h
    "MetalRenderContext_new": (c_void_p, [c_int]),
    "MetalRenderContext_release": (None, [c_void_p]),
    "MetalRenderContext_setupMetal": (c_int, [c_void_p]),
    "MetalRenderContext_makeBGRATexture": (PTX, [c_void_p, c_int, c_int, c_void_p, c_int]),
    "MBDTextureRelease": (None, [PTX]),
    "MetalRenderContext_waitUntilCompleted": (c_int, [c_void_p]),
    "MetalRenderContext_ctx": (c_void_p, [c_void_p]),
    "MetalRenderContext_launchCount": (ctypes.c_ulonglong, [c_void_p]),
    "MetalBT709Decoder_new": (c_void_p, []),
    "MetalBT709Decoder_release": (None, [c_void_p]),
    "MetalBT709Decoder_setMetalRenderContext": (None, [c_void_p, c_void_p]),
    "MetalBT709Decoder_setGamma": (None, [c_void_p, c_int]),
    "MetalBT709Decoder_gamma": (c_int, [c_void_p]),
    "MetalBT709Decoder_setUseComputeRenderer": (None, [c_void_p, c_int]),
    "MetalBT709Decoder_setHasAlphaChannel": (None, [c_void_p, c_int]),
    "MetalBT709Decoder_setupMetal": (c_int, [c_void_p]),
    "MetalBT709Decoder_decodeBT709": (c_int, [c_void_p, PPB, PPB, PTX, c_void_p, c_int, c_int, c_int]),
    "MetalScaleRenderContext_new": (c_void_p, []),
    "MetalScaleRenderContext_release": (None, [c_void_p]),
    "MetalScaleRenderContext_renderScaled": (c_int, [c_void_p, c_void_p, PTX, c_int, c_int, c_void_p, PTX, c_int]),
    "BGRAToBT709Converter_createCoreVideoYCbCrBuffer": (PPB, [c_int, c_int]),
    "MBDPixelBufferRelease": (None, [PPB]),
    "BGRAToBT709Converter_setBT709Attributes": (c_int, [PPB]),
    "BGRAToBT709Converter_convert": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
    "BGRAToBT709Converter_unconvert": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
    "BGRAToBT709Converter_copyBT709ToCoreVideo": (c_int, [c_void_p, c_void_p, PPB]),
    "BGRAToBT709Converter_createYCbCrFromBGRA": (PPB, [c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    "BGRAToBT709Converter_copyYCBCr": (c_int, [c_void_p, PPB, c_void_p, c_void_p, c_void_p]),
    "cvpbu_ycbcr_subsample": (c_int, [c_void_p, c_void_p, c_int, c_int, PPB, c_int, c_int]),
    "AAPLRenderer_new": (c_void_p, [c_void_p, c_int, c_int]),
    "AAPLRenderer_release": (None, [c_void_p]),
    "AAPLRenderer_drawFrame": (c_int, [c_void_p, PPB, PPB, PTX, c_int]),
}


def lib():
    """Load (building first if the sources are newer) and return the CDLL with prototypes set."""
    global _lib
    if _lib is not None:
        return _lib
    if "BT709_B200_LIB" not in os.environ and _needs_build():
        build()
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(L, name)  # AttributeError = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
