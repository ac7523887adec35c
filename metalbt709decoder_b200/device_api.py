"""Device-pointer API: include/bt709_b200.h over torch CUDA tensors.

torch is used only as the owner of device memory and streams (tensor.data_ptr(),
torch.cuda.current_stream()); every computation happens in libbt709_b200.so.
"""
import ctypes

from . import _lib
from ._lib import Plane

GAMMA_APPLE, GAMMA_SRGB, GAMMA_LINEAR, GAMMA_BT709 = 0, 1, 2, 3
GAMMA_NAMES = {"apple": 0, "srgb": 1, "linear": 2, "bt709": 3}

OK, E_NULL, E_DIMS, E_PITCH, E_GAMMA, E_ALPHA_MODE, E_CUDA, E_SCRATCH, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6, -7, -8


class BT709Error(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = _lib.lib().bt709_status_string(status).decode()
        super().__init__(f"{where}: {msg} ({status})")


def plane_of(t):
    """bt709_plane for a (frames, rows, row_bytes...) uint8/int32 CUDA tensor whose rows are contiguous."""
    assert t.is_cuda, "device tensors only: the hot path has no CPU fallback"
    es = t.element_size()
    if t.dim() == 2:
        return Plane(t.data_ptr(), t.stride(0) * es, 0)
    assert t.dim() >= 3
    return Plane(t.data_ptr(), t.stride(1) * es, t.stride(0) * es)


def _stream_ptr(stream):
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return ctypes.c_void_p(stream.cuda_stream)


class BT709Context:
    """bt709_ctx: one per device.  Replaces MetalRenderContext + -setupMetal."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        h = ctypes.c_void_p()
        rc = self._L.bt709_ctx_create(int(device), ctypes.byref(h))
        if rc != OK:
            raise BT709Error(rc, "bt709_ctx_create")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.bt709_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def launch_count(self):
        return int(self._L.bt709_launch_count(self._h))

    def last_cuda_error(self):
        return int(self._L.bt709_last_cuda_error(self._h))

    # ---- raw status-returning calls (tests exercise the error codes through these) ----
    def decode_nv12_raw(self, gamma, y, cbcr, alpha, out, w, h, nframes, stream=None, wait=False):
        return self._L.bt709_decode_nv12(self._h, gamma, y, cbcr, alpha, out, w, h, nframes, _stream_ptr(stream),
                                         int(wait))

    # ---- tensor conveniences ----
    def decode_nv12(self, gamma, y, cbcr, out, alpha_y=None, w=None, h=None, stream=None, wait=False):
        """y (F,H,Wp) u8, cbcr (F,ceil(H/2),Cp) u8, out (F,H,Wo,4) u8 device tensors."""
        F, H = y.shape[0], y.shape[1]
        w = w if w is not None else out.shape[2]
        h = h if h is not None else H
        py, pc, po = plane_of(y), plane_of(cbcr), plane_of(out)
        pa = plane_of(alpha_y) if alpha_y is not None else None
        rc = self._L.bt709_decode_nv12(self._h, gamma, ctypes.byref(py), ctypes.byref(pc),
                                       ctypes.byref(pa) if pa is not None else None, ctypes.byref(po), w, h, F,
                                       _stream_ptr(stream), int(wait))
        if rc != OK:
            raise BT709Error(rc, "bt709_decode_nv12")

    def rescale_bgra(self, src, dst, stream=None, wait=False):
        """src (F,IH,IW,4) -> dst (F,OH,OW,4), u8 device tensors."""
        F, ih, iw = src.shape[0], src.shape[1], src.shape[2]
        oh, ow = dst.shape[1], dst.shape[2]
        pi, po = plane_of(src), plane_of(dst)
        rc = self._L.bt709_rescale_bgra(self._h, ctypes.byref(pi), iw, ih, ctypes.byref(po), ow, oh, F,
                                        _stream_ptr(stream), int(wait))
        if rc != OK:
            raise BT709Error(rc, "bt709_rescale_bgra")

    def decode_rescale_nv12(self, gamma, y, cbcr, scratch, out, alpha_y=None, w=None, h=None, stream=None, wait=False):
        F, H = y.shape[0], y.shape[1]
        h = h if h is not None else H
        w = w if w is not None else (scratch.shape[2] if scratch is not None else y.shape[2])
        oh, ow = out.shape[1], out.shape[2]
        py, pc, po = plane_of(y), plane_of(cbcr), plane_of(out)
        ps = plane_of(scratch) if scratch is not None else None
        pa = plane_of(alpha_y) if alpha_y is not None else None
        rc = self._L.bt709_decode_rescale_nv12(self._h, gamma, ctypes.byref(py), ctypes.byref(pc),
                                               ctypes.byref(pa) if pa is not None else None,
                                               ctypes.byref(ps) if ps is not None else None, ctypes.byref(po),
                                               w, h, ow, oh, F, _stream_ptr(stream), int(wait))
        if rc != OK:
            raise BT709Error(rc, "bt709_decode_rescale_nv12")

    def decode_rescale_nv12_fused(self, gamma, y, cbcr, scratch, out, w=None, h=None, stream=None, wait=False):
        """One kernel, the intermediate stays in L2 (bt709_decode_rescale_nv12_fused); BT709Error(-8) if it does not apply."""
        F, H = y.shape[0], y.shape[1]
        h = h if h is not None else H
        w = w if w is not None else y.shape[2]
        oh, ow = out.shape[1], out.shape[2]
        py, pc, po, ps = plane_of(y), plane_of(cbcr), plane_of(out), plane_of(scratch)
        rc = self._L.bt709_decode_rescale_nv12_fused(self._h, gamma, ctypes.byref(py), ctypes.byref(pc), ctypes.byref(ps),
                                                     ctypes.byref(po), w, h, ow, oh, F, _stream_ptr(stream), int(wait))
        if rc != OK:
            raise BT709Error(rc, "bt709_decode_rescale_nv12_fused")

    def encode_bgra_to_nv12(self, gamma_in, gamma_out, bgra, y, cbcr, stream=None, wait=False):
        """bgra (F,H,W,4) -> y (F,H,Wp), cbcr (F,H/2,Cp)."""
        F, h, w = bgra.shape[0], bgra.shape[1], bgra.shape[2]
        pi, py, pc = plane_of(bgra), plane_of(y), plane_of(cbcr)
        rc = self._L.bt709_encode_bgra_to_nv12(self._h, gamma_in, gamma_out, ctypes.byref(pi), ctypes.byref(py),
                                               ctypes.byref(pc), w, h, F, _stream_ptr(stream), int(wait))
        if rc != OK:
            raise BT709Error(rc, "bt709_encode_bgra_to_nv12")

    def convert_packed444(self, bgra_words, out_words, stream=None, wait=False, gamma=GAMMA_APPLE):
        """(h,w) int32 BGRA words -> Cr<<16 | Cb<<8 | Y; gamma picks the per-pixel encoder (BT709.h:610-944)."""
        h, w = bgra_words.shape
        rc = self._L.bt709_convert_packed444_gamma(self._h, int(gamma), bgra_words.data_ptr(), out_words.data_ptr(), w, h,
                                                   _stream_ptr(stream), int(wait))
        if rc != OK:
            raise BT709Error(rc, "bt709_convert_packed444_gamma")

    def unconvert_packed444(self, ycbcr_words, out_words, stream=None, wait=False, gamma=GAMMA_APPLE):
        h, w = ycbcr_words.shape
        rc = self._L.bt709_unconvert_packed444_gamma(self._h, int(gamma), ycbcr_words.data_ptr(), out_words.data_ptr(),
                                                     w, h, _stream_ptr(stream), int(wait))
        if rc != OK:
            raise BT709Error(rc, "bt709_unconvert_packed444_gamma")

    def copy_packed444_to_nv12(self, ycbcr_words, y, cbcr, stream=None, wait=False):
        h, w = ycbcr_words.shape
        py, pc = plane_of(y), plane_of(cbcr)
        rc = self._L.bt709_copy_packed444_to_nv12(self._h, ycbcr_words.data_ptr(), ctypes.byref(py), ctypes.byref(pc),
                                                  w, h, _stream_ptr(stream), int(wait))
        if rc != OK:
            raise BT709Error(rc, "bt709_copy_packed444_to_nv12")
