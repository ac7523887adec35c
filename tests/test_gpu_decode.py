"""Parity of the CUDA decode (pass 1) with the CPU oracle, through the C ABI.  -m gpu."""
import ctypes

import numpy as np
import pytest
import torch

from _oracle import Oracle, frame_lattice, frame_random

pytestmark = pytest.mark.gpu
ORA = Oracle()
GAMMAS = [0, 1, 2, 3]


def _check(got, want, what):
    from _gpu import parity_report
    exact, mx = parity_report(got, want)
    # north_star gate: +-1 LSB per channel; this implementation is designed to be bit-exact
    assert mx <= 1, f"{what}: max abs diff {mx} > 1 LSB"
    assert exact == 1.0, f"{what}: bit-exact fraction {exact:.6f} (max diff {mx})"


@pytest.mark.parametrize("gamma", GAMMAS)
def test_all_2_24_inputs_bit_exact(gamma):
    """F-lattice: every (Y,Cb,Cr) triple in one 4096x4096 frame."""
    from _gpu import decode
    y, c = frame_lattice()
    _check(decode(gamma, y, c), ORA.decode_nv12(gamma, y, c), f"lattice gamma={gamma}")


@pytest.mark.parametrize("gamma", GAMMAS)
@pytest.mark.parametrize("w,h", [(1920, 1080), (256, 2), (4, 2), (132, 6)])
def test_random_frames_fast_path(gamma, w, h):
    from _gpu import decode
    y, c = frame_random(w, h, seed=709)
    _check(decode(gamma, y, c), ORA.decode_nv12(gamma, y, c), f"{w}x{h}")


@pytest.mark.parametrize("w,h", [(7680, 8), (4224, 6), (8196, 4), (4096, 4), (5000, 3)])
def test_rows_wider_than_32_tiles(w, h):
    """more than 32 tiles of 128 pixels per row: the CTA's warps loop over the tile columns.  (The ring kernel
    gives each warp ONE tile column and once decoded only the first 4096 pixels of such rows.)"""
    from _gpu import decode
    y, c = frame_random(w, h, seed=712)
    for gamma in (0, 2):
        _check(decode(gamma, y, c), ORA.decode_nv12(gamma, y, c), f"{w}x{h} gamma={gamma}")


@pytest.mark.parametrize("gamma", GAMMAS)
@pytest.mark.parametrize("w,h", [(37, 21), (1, 1), (2, 2), (3, 5), (130, 7), (258, 3)])
def test_ragged_sizes_generic_path(gamma, w, h):
    """odd widths / heights: chroma plane is ceil(w/2) x ceil(h/2) pairs"""
    from _gpu import decode
    y, c = frame_random(w, h, seed=710, full_range=True)
    _check(decode(gamma, y, c), ORA.decode_nv12(gamma, y, c), f"{w}x{h}")


def test_full_range_bytes_saturate_like_the_reference():
    from _gpu import decode
    y, c = frame_random(512, 64, seed=710, full_range=True)
    for gamma in GAMMAS:
        _check(decode(gamma, y, c), ORA.decode_nv12(gamma, y, c), "full range")


def test_odd_height_on_fast_path():
    from _gpu import decode
    y, c = frame_random(256, 5, seed=3)
    _check(decode(0, y, c), ORA.decode_nv12(0, y, c), "256x5")


def test_padded_pitches_and_unaligned_bases():
    """planes with row padding (CoreVideo bytesPerRow) and a base pointer that is not 4-aligned"""
    from _gpu import ctx, dev
    w, h = 200, 10
    y, c = frame_random(w, h, seed=11)
    want = ORA.decode_nv12(3, y, c)
    yp = np.zeros((h, 256), np.uint8); yp[:, :w] = y
    cp = np.zeros((h // 2, 320), np.uint8); cp[:, :w] = c
    ty, tc = dev(yp)[None], dev(cp)[None]
    out = torch.zeros((1, h, 208, 4), dtype=torch.uint8, device="cuda")
    ctx().decode_nv12(3, ty, tc, out, w=w, h=h, wait=True)
    _check(out[0, :, :w].cpu().numpy(), want, "padded")
    assert int(out[0, :, w:].max()) == 0  # padding untouched
    # unaligned base: shift the Y plane by one byte inside a larger allocation
    flat = torch.zeros(h * 256 + 1, dtype=torch.uint8, device="cuda")
    flat[1:] = ty.reshape(-1)
    ty2 = flat[1:].view(1, h, 256)
    out.zero_()
    ctx().decode_nv12(3, ty2, tc, out, w=w, h=h, wait=True)
    _check(out[0, :, :w].cpu().numpy(), want, "unaligned")


@pytest.mark.parametrize("w,h,alpha", [(854, 480, False), (1366, 768, False), (130, 6, True), (6, 2, False)])
def test_even_widths_not_multiple_of_4_on_the_fast_path(w, h, alpha):
    """w % 4 == 2 with 4-byte-aligned pitches (what CoreVideo hands over for 854x480 or 1366x768): the last
    lane of a row owns two pixels only; the two pixels after the row end must stay untouched"""
    from _gpu import ctx, dev
    y, c = frame_random(w, h, seed=21, full_range=True)
    F = 3
    pw = (w + 3) // 4 * 4 + 4
    ty = torch.full((F, h, pw), 0x11, dtype=torch.uint8, device="cuda")
    tc = torch.full((F, h // 2, pw), 0x80, dtype=torch.uint8, device="cuda")
    ty[:, :, :w] = dev(y)[None]
    tc[:, :, :w] = dev(c)[None]
    ta = None
    if alpha:
        ay, _ = frame_random(w, h, seed=22, full_range=True)
        ta = torch.full((F, h, pw), 0x22, dtype=torch.uint8, device="cuda")
        ta[:, :, :w] = dev(ay)[None]
    out = torch.full((F, h, pw, 4), 0xEE, dtype=torch.uint8, device="cuda")
    gamma = 1 if alpha else 3
    ctx().decode_nv12(gamma, ty[:, :, :w], tc[:, :, :w], out[:, :, :w], alpha_y=None if ta is None else ta[:, :, :w], wait=True)
    want = ORA.decode_nv12(gamma, y, c, alpha_y=ay if alpha else None)
    got = out.cpu().numpy()
    for f in range(F):
        _check(got[f, :, :w], want, f"{w}x{h} frame {f}")
    assert (got[:, :, w:] == 0xEE).all(), "wrote past the end of a row"


@pytest.mark.parametrize("w,h,F,alpha", [(384, 6, 7, False), (1920, 4, 3, False), (3840, 2, 2, False), (640, 10, 4, True),
                                         (1152, 6, 5, False)])
def test_linear_tile_order_on_contiguous_batches(w, h, F, alpha):
    """rows of 3, 5, 9, 15 or 30 tiles do not divide the CTA's 32 warps: contiguous batches are then decoded in
    linear tile order (the tile column steps and wraps into the next row pair, across frame seams too)"""
    from _gpu import ctx, dev
    frames = [frame_random(w, h, seed=300 + f) for f in range(F)]
    ty = torch.stack([dev(y) for y, _ in frames])
    tc = torch.stack([dev(c) for _, c in frames])
    ta = None
    if alpha:
        alphas = [frame_random(w, h, seed=400 + f, full_range=True)[0] for f in range(F)]
        ta = torch.stack([dev(a) for a in alphas])
    out = torch.zeros((F, h, w, 4), dtype=torch.uint8, device="cuda")
    gamma = 1 if alpha else 0
    ctx().decode_nv12(gamma, ty, tc, out, alpha_y=ta, wait=True)
    got = out.cpu().numpy()
    for f in range(F):
        want = ORA.decode_nv12(gamma, frames[f][0], frames[f][1], alpha_y=alphas[f] if alpha else None)
        _check(got[f], want, f"{w}x{h} frame {f}")


def test_batch_of_frames_matches_per_frame():
    from _gpu import ctx, dev
    w, h, F = 640, 360, 5
    ys, cs, wants = [], [], []
    for f in range(F):
        y, c = frame_random(w, h, seed=709 + f)
        ys.append(y); cs.append(c); wants.append(ORA.decode_nv12(0, y, c))
    ty, tc = dev(np.stack(ys)), dev(np.stack(cs))
    out = torch.zeros((F, h, w, 4), dtype=torch.uint8, device="cuda")
    ctx().decode_nv12(0, ty, tc, out, wait=True)
    _check(out.cpu().numpy(), np.stack(wants), "batch")


@pytest.mark.parametrize("w,h", [(2048, 16), (50, 9)])
def test_alpha_stream(w, h):
    """C4: RGB stream + alpha stream's Y plane -> BGRA with A, sRGB gamma, no premultiply"""
    from _gpu import decode
    y, c = frame_random(w, h, seed=21)
    a, _ = frame_random(w, h, seed=22, full_range=True)
    got = decode(1, y, c, alpha=a)
    want = ORA.decode_nv12(1, y, c, alpha_y=a)
    _check(got, want, "alpha")
    assert set(np.unique(want[..., 3])) != {255}


def test_error_codes():
    from _gpu import ctx, dev
    from metalbt709decoder_b200 import BT709Error
    from metalbt709decoder_b200._lib import Plane
    C = ctx()
    y, c = frame_random(16, 8)
    ty, tc = dev(y)[None], dev(c)[None]
    out = torch.zeros((1, 8, 16, 4), dtype=torch.uint8, device="cuda")
    with pytest.raises(BT709Error) as e:
        C.decode_nv12(7, ty, tc, out)
    assert e.value.status == -4  # BT709_E_GAMMA
    with pytest.raises(BT709Error) as e:
        C.decode_nv12(0, ty, tc, out, alpha_y=ty)
    assert e.value.status == -5  # alpha needs sRGB (MetalBT709Decoder.m:165-169)
    with pytest.raises(BT709Error) as e:
        C.decode_nv12(0, ty, tc, out, w=0)
    assert e.value.status == -2
    with pytest.raises(BT709Error) as e:
        C.decode_nv12(0, ty, tc, out, w=32)  # pitch (16) < row bytes (32)
    assert e.value.status == -3
    p = Plane(None, 16, 0)
    assert C.decode_nv12_raw(0, ctypes.byref(p), ctypes.byref(p), None, ctypes.byref(p), 16, 8, 1) == -1
    assert C.last_cuda_error() == 0


def test_async_then_stream_sync_and_launch_count():
    from _gpu import ctx, dev
    C = ctx()
    y, c = frame_random(1920, 1080, seed=1)
    ty, tc = dev(y)[None], dev(c)[None]
    out = torch.zeros((1, 1080, 1920, 4), dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    n0 = C.launch_count
    with torch.cuda.stream(s):
        C.decode_nv12(3, ty, tc, out, stream=s, wait=False)
    s.synchronize()
    assert C.launch_count == n0 + 1
    _check(out[0].cpu().numpy(), ORA.decode_nv12(3, y, c), "async")


def test_uhd_frame_size_independent_properties():
    """BASELINE config sizes: properties that do not need the oracle at full size.
    (1) decode is a per-pixel map: shuffling 2x2 blocks commutes with decoding;
    (2) neutral chroma gives R=G=B; (3) a spot-checked band matches the oracle."""
    from _gpu import decode
    w, h = 3840, 2160
    y, c = frame_random(w, h, seed=4)
    got = decode(0, y, c)
    band = slice(1000, 1016)
    _check(got[band], ORA.decode_nv12(0, y[band], c[500:508]), "UHD band")
    yb = y.reshape(h // 2, 2, w // 2, 2)
    cb = c.reshape(h // 2, w // 2, 2)
    perm = np.random.default_rng(0).permutation(w // 2)
    y2 = yb[:, :, perm, :].reshape(h, w)
    c2 = cb[:, perm, :].reshape(h // 2, w)
    got2 = decode(0, y2, c2)
    ref2 = got.reshape(h // 2, 2, w // 2, 2, 4)[:, :, perm].reshape(h, w, 4)
    assert np.array_equal(got2, ref2)
    cn = np.full_like(c, 128)
    g = decode(3, y, cn)
    assert np.array_equal(g[..., 0], g[..., 1]) and np.array_equal(g[..., 1], g[..., 2])
    assert (g[..., 3] == 255).all()


def test_bulk_staged_variant_matches(tmp_path):
    """The cp.async.bulk (TMA 1-D) shared-memory-ring variant of the decode kernel is opt-in
    (BT709_BULK=1, read once per process): run it in a subprocess and compare with the oracle."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests")
from _oracle import Oracle, frame_random
from _gpu import decode
ora = Oracle()
for (w, h, g) in [(1920, 64, 3), (3840, 8, 0), (2048, 16, 2), (384, 6, 1)]:
    y, c = frame_random(w, h, seed=5)
    assert np.array_equal(decode(g, y, c), ora.decode_nv12(g, y, c)), (w, h, g)
print("bulk ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BT709_BULK="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "bulk ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("switch", ["BT709_NO_RING", "BT709_NO_LINEAR", "BT709_NO_COLS2", "BT709_NO_ELOG", "BT709_NO_PAIR",
                                    "BT709_NO_BLOG"])
def test_superseded_kernel_variants_still_match(switch):
    """the variants kept for reproduction (band layout, register-prefetch decode, one-column
    rescale, uniform-cell encode LUT) are selected per process by an environment variable; each must stay exact"""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "tests")
from _oracle import Oracle, frame_random, bgra_random
from _gpu import decode, ctx, dev
ora = Oracle()
for (w, h, g) in [(1920, 8, 3), (3840, 4, 2), (640, 6, 0), (2048, 4, 1)]:
    y, c = frame_random(w, h, seed=6)
    assert np.array_equal(decode(g, y, c), ora.decode_nv12(g, y, c)), (w, h, g)
for (iw, ih, ow, oh) in [(192, 108, 256, 144), (300, 200, 150, 100), (64, 64, 64, 64)]:
    src = bgra_random(iw, ih, seed=7, alpha=None)
    out = torch.zeros((1, oh, ow, 4), dtype=torch.uint8, device="cuda")
    ctx().rescale_bgra(dev(src)[None], out, wait=True)
    assert np.array_equal(out[0].cpu().numpy(), ora.rescale_bgra(src, ow, oh)), (iw, ih, ow, oh)
print("variant ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **{switch: "1"})
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout + r.stderr


def test_padded_frame_strides_take_the_wrap_path():
    """frames separated by padding (frame_stride > rows * pitch): the kernels step across frame seams"""
    from _gpu import ctx
    w, h, F = 256, 36, 7
    ys, cs, wants = [], [], []
    ybig = torch.zeros((F, h + 3, w + 64), dtype=torch.uint8, device="cuda")
    cbig = torch.zeros((F, h // 2 + 5, w + 128), dtype=torch.uint8, device="cuda")
    obig = torch.zeros((F, h + 2, w + 16, 4), dtype=torch.uint8, device="cuda")
    for f in range(F):
        y, c = frame_random(w, h, seed=40 + f)
        ybig[f, :h, :w] = torch.from_numpy(y).cuda()
        cbig[f, :h // 2, :w] = torch.from_numpy(c).cuda()
        wants.append(ORA.decode_nv12(3, y, c))
    for env_gamma in (3, 2):
        obig.zero_()
        ctx().decode_nv12(env_gamma, ybig[:, :h, :w], cbig[:, :h // 2, :w], obig[:, :h, :w], w=w, h=h, wait=True)
        got = obig[:, :h, :w].cpu().numpy()
        if env_gamma == 3:
            _check(got, np.stack(wants), "strided batch")
        assert int(obig[:, h:].max()) == 0 and int(obig[:, :, w:].max()) == 0  # padding untouched


def test_offsets_beyond_32_bits_are_split_across_launches():
    """a batch whose planes span more than 4 GiB (sparse frame strides): the 32-bit unit offsets of the
    fast kernels cannot address it in one launch, so the launcher splits it"""
    from _gpu import ctx
    from metalbt709decoder_b200._lib import Plane
    import ctypes
    w, h, F = 256, 8, 3
    stride_y = (5 << 30) // 2 + 4096          # 2.5 GiB between Y frames
    stride_o = (36 << 30) + 65536             # 36 GiB between BGRA frames: > 2^32 * 16 B in two frames
    free, _ = torch.cuda.mem_get_info()
    need = stride_y * (F - 1) + stride_o * (F - 1) + (1 << 30)
    if free < need:
        pytest.skip("not enough free device memory")
    ybuf = torch.empty(stride_y * (F - 1) + w * h, dtype=torch.uint8, device="cuda")
    obuf = torch.empty(stride_o * (F - 1) + w * h * 4, dtype=torch.uint8, device="cuda")
    cbuf = torch.empty((F, h // 2, w), dtype=torch.uint8, device="cuda")
    wants = []
    for f in range(F):
        y, c = frame_random(w, h, seed=60 + f)
        ybuf[f * stride_y:f * stride_y + w * h] = torch.from_numpy(y).cuda().reshape(-1)
        cbuf[f] = torch.from_numpy(c).cuda()
        wants.append(ORA.decode_nv12(0, y, c))
    C = ctx()
    py = Plane(ybuf.data_ptr(), w, stride_y)
    pc = Plane(cbuf.data_ptr(), w, (h // 2) * w)
    po = Plane(obuf.data_ptr(), w * 4, stride_o)
    rc = C.decode_nv12_raw(0, ctypes.byref(py), ctypes.byref(pc), None, ctypes.byref(po), w, h, F, wait=True)
    assert rc == 0
    for f in range(F):
        got = obuf[f * stride_o:f * stride_o + w * h * 4].view(h, w, 4).cpu().numpy()
        _check(got, wants[f], f"frame {f} of a >4 GiB batch")
