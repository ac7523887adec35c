"""ctypes bindings for the CPU oracle (oracle/) — the checker used by tests/.

The oracle is test infrastructure: nothing in metalbt709decoder_b200/ imports this.
"""
import ctypes
import os
import subprocess
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

GAMMA_APPLE, GAMMA_SRGB, GAMMA_LINEAR, GAMMA_BT709 = 0, 1, 2, 3
GAMMA_NAMES = {0: "apple", 1: "srgb", 2: "linear", 3: "bt709"}

_c_int_p = ctypes.POINTER(ctypes.c_int)


def _build():
    subprocess.run(["make", "-s", "-C", ODIR], check=True, stdout=subprocess.DEVNULL)


def _load(path):
    return ctypes.CDLL(path)


def load_oracle():
    so = os.path.join(ODIR, "liboracle_bt709.so")
    src = os.path.join(ODIR, "oracle_bt709.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        _build()
    lib = _load(so)
    lib.oracle_srgb_to_linear.restype = ctypes.c_float
    lib.oracle_srgb_to_linear.argtypes = [ctypes.c_float]
    lib.oracle_linear_to_srgb.restype = ctypes.c_float
    lib.oracle_linear_to_srgb.argtypes = [ctypes.c_float]
    return lib


def load_ref():
    """The reference headers compiled where they lie (oracle/_ref). None if unavailable."""
    so = os.path.join(ODIR, "_ref", "libref_bt709.so")
    if not os.path.exists(so):
        if os.path.exists("/root/reference/Renderer/BT709.h"):
            _build()
        if not os.path.exists(so):
            return None
    lib = _load(so)
    lib.ref_srgb_to_linear.restype = ctypes.c_float
    lib.ref_srgb_to_linear.argtypes = [ctypes.c_float]
    lib.ref_linear_to_srgb.restype = ctypes.c_float
    lib.ref_linear_to_srgb.argtypes = [ctypes.c_float]
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _L(v):
    return ctypes.c_long(int(v))


class Oracle:
    """Thin numpy-facing wrapper.  prefix = 'oracle_' (restatement) or 'ref_' (reference)."""

    def __init__(self, lib=None, prefix="oracle_"):
        self.lib = lib if lib is not None else load_oracle()
        self.prefix = prefix

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- per pixel ----
    def decode_pixel(self, gamma, y, cb, cr):
        r, g, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._f("decode_pixel")(gamma, y, cb, cr, ctypes.byref(r), ctypes.byref(g), ctypes.byref(b))
        return r.value, g.value, b.value

    def encode444_pixel(self, gamma, r, g, b):
        y, cb, cr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._f("encode444_pixel")(gamma, r, g, b, ctypes.byref(y), ctypes.byref(cb), ctypes.byref(cr))
        return y.value, cb.value, cr.value

    def linear_rgb_to_ycbcr(self, r, g, b, apply_gamma):
        y, cb, cr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._f("linear_rgb_to_ycbcr")(r, g, b, apply_gamma, ctypes.byref(y), ctypes.byref(cb), ctypes.byref(cr))
        return y.value, cb.value, cr.value

    def ycbcr_to_linear_rgb(self, y, cb, cr, apply_gamma):
        r, g, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._f("ycbcr_to_linear_rgb")(y, cb, cr, apply_gamma, ctypes.byref(r), ctypes.byref(g), ctypes.byref(b))
        return r.value, g.value, b.value

    def average_block(self, rgb12, gin, gout):
        arr = (ctypes.c_int * 12)(*rgb12)
        y4 = (ctypes.c_int * 4)()
        cb, cr = ctypes.c_int(), ctypes.c_int()
        self._f("average_block")(arr, gin, gout, y4, ctypes.byref(cb), ctypes.byref(cr))
        return list(y4), cb.value, cr.value

    def srgb_average_cbcr(self, rgb12):
        arr = (ctypes.c_int * 12)(*rgb12)
        self._f("srgb_average_cbcr")(arr)
        return list(arr)

    def alpha_pixel(self, a):
        return self._f("alpha_pixel")(a)

    def roundtrip_histogram(self, gamma):
        b = (ctypes.c_longlong * 11)()
        self._f("roundtrip_histogram")(gamma, b)
        return list(b)

    # ---- frames ----
    def decode_nv12(self, gamma, y, cbcr, alpha_y=None, nthreads=8):
        """y: (h, py) uint8; cbcr: (ceil(h/2), pc) uint8 -> (h, w, 4) uint8, w = y.shape[1]"""
        h, w = y.shape
        out = np.zeros((h, w, 4), np.uint8)
        if self.prefix == "ref_":
            assert alpha_y is None
            self.lib.ref_decode_nv12(gamma, _p(y), _L(y.strides[0]), _p(cbcr), _L(cbcr.strides[0]),
                                     _p(out), _L(out.strides[0]), w, h)
            return out
        rc = self.lib.oracle_decode_nv12(
            gamma, _p(y), _L(y.strides[0]), _p(cbcr), _L(cbcr.strides[0]),
            _p(alpha_y) if alpha_y is not None else None,
            _L(alpha_y.strides[0] if alpha_y is not None else 0),
            _p(out), _L(out.strides[0]), w, h, nthreads)
        assert rc == 0, rc
        return out

    def encode_bgra_nv12(self, gin, gout, bgra, nthreads=8):
        """bgra: (h, w, 4) uint8 -> y (h, w), cbcr (h/2, w)"""
        h, w, _ = bgra.shape
        y = np.zeros((h, w), np.uint8)
        c = np.zeros((h // 2, w), np.uint8)
        if self.prefix == "ref_":
            self.lib.ref_encode_bgra_nv12(gin, gout, _p(bgra), _L(bgra.strides[0]), _p(y), _L(w), _p(c), _L(w), w, h)
            return y, c
        rc = self.lib.oracle_encode_bgra_nv12(gin, gout, _p(bgra), _L(bgra.strides[0]), _p(y), _L(w),
                                              _p(c), _L(w), w, h, nthreads)
        assert rc == 0, rc
        return y, c

    def rescale_bgra(self, src, ow, oh, nthreads=8):
        ih, iw, _ = src.shape
        out = np.zeros((oh, ow, 4), np.uint8)
        rc = self.lib.oracle_rescale_bgra(_p(src), _L(src.strides[0]), iw, ih, _p(out), _L(out.strides[0]),
                                          ow, oh, nthreads)
        assert rc == 0, rc
        return out

    def convert_packed444(self, bgra_words):
        h, w = bgra_words.shape
        out = np.zeros((h, w), np.uint32)
        rc = self.lib.oracle_convert_packed444(_p(bgra_words), _p(out), w, h)
        return rc, out

    def unconvert_packed444(self, ycbcr_words):
        h, w = ycbcr_words.shape
        out = np.zeros((h, w), np.uint32)
        rc = self.lib.oracle_unconvert_packed444(_p(ycbcr_words), _p(out), w, h)
        return rc, out

    def copy_packed444_to_nv12(self, ycbcr_words):
        h, w = ycbcr_words.shape
        y = np.zeros((h, w), np.uint8)
        c = np.zeros(((h + 1) // 2, w + (w & 1)), np.uint8)
        self.lib.oracle_copy_packed444_to_nv12(_p(ycbcr_words), _p(y), _L(w), _p(c), _L(c.strides[0]), w, h)
        return y, c


def ref_oracle():
    lib = load_ref()
    return Oracle(lib, "ref_") if lib is not None else None


# ---------------------------------------------------------------- synthetic frames
def frame_random(w, h, seed=709, full_range=False):
    """F-random / F-full: NV12 planes. Returns y (h, w), cbcr (ceil(h/2), 2*ceil(w/2))."""
    rng = np.random.default_rng(seed)
    ch, cw = (h + 1) // 2, (w + 1) // 2
    if full_range:
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        c = rng.integers(0, 256, (ch, cw * 2), dtype=np.uint8)
    else:
        y = rng.integers(16, 236, (h, w), dtype=np.uint8)
        c = rng.integers(16, 241, (ch, cw * 2), dtype=np.uint8)
    return y, c


def frame_lattice():
    """F-lattice: every (Y, Cb, Cr) in one 4096x4096 frame (chroma 2048x2048 pairs)."""
    idx = np.arange(2048 * 2048, dtype=np.uint32).reshape(2048, 2048)
    pair = idx % 65536
    k = (idx // 65536).astype(np.uint8)
    c = np.empty((2048, 2048, 2), np.uint8)
    c[..., 0] = pair & 255
    c[..., 1] = pair >> 8
    y = np.empty((4096, 4096), np.uint8)
    y[0::2, 0::2] = 4 * k
    y[0::2, 1::2] = 4 * k + 1
    y[1::2, 0::2] = 4 * k + 2
    y[1::2, 1::2] = 4 * k + 3
    return y, c.reshape(2048, 4096)


def frame_ramp(w, h):
    """F-ramp: the gamma_write pattern, gray = (col/2) & 255 in 2x2 blocks (gamma_write.m:280-298)."""
    g = ((np.arange(w) // 2) & 255).astype(np.uint8)
    bgra = np.empty((h, w, 4), np.uint8)
    bgra[..., 0] = g
    bgra[..., 1] = g
    bgra[..., 2] = g
    bgra[..., 3] = 255
    return bgra


def bgra_random(w, h, seed=711, alpha=255):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if alpha is not None:
        a[..., 3] = alpha
    return a
