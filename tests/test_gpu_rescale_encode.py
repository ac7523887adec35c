"""Parity of pass 2 (rescale), the 4:2:0 encode and the packed 4:4:4 harness path.  -m gpu."""
import numpy as np
import pytest
import torch

from _oracle import GAMMA_APPLE, GAMMA_LINEAR, GAMMA_SRGB, Oracle, bgra_random, frame_ramp, frame_random

pytestmark = pytest.mark.gpu
ORA = Oracle()


def _exact(got, want, what):
    from _gpu import parity_report
    exact, mx = parity_report(got, want)
    assert mx <= 1, f"{what}: max abs diff {mx}"
    assert exact == 1.0, f"{what}: bit-exact fraction {exact:.6f}"


# ------------------------------------------------------------------------------ pass 2
@pytest.mark.parametrize("iw,ih,ow,oh", [(1920, 1080, 2560, 1440), (64, 48, 64, 48), (64, 48, 31, 17), (5, 3, 40, 33),
                                         (320, 240, 1, 1), (2, 2, 257, 3), (640, 360, 320, 180), (1000, 300, 77, 101),
                                         (200, 301, 300, 100), (128, 400, 128, 99), (3840, 2160, 1280, 720)])
def test_rescale_matches_oracle(iw, ih, ow, oh):
    """PARITY UNPINNED in the reference (no CPU twin); pinned to the restatement."""
    from _gpu import ctx, dev
    src = bgra_random(iw, ih, seed=5, alpha=None)
    want = ORA.rescale_bgra(src, ow, oh)
    out = torch.zeros((1, oh, ow, 4), dtype=torch.uint8, device="cuda")
    ctx().rescale_bgra(dev(src)[None], out, wait=True)
    _exact(out[0].cpu().numpy(), want, f"rescale {iw}x{ih}->{ow}x{oh}")


def test_rescale_identity_size_is_identity():
    """same-size bilinear with pixel-centre sampling reproduces the source (sRGB round trip is exact)"""
    from _gpu import ctx, dev
    src = bgra_random(128, 64, seed=6, alpha=None)
    out = torch.zeros((1, 64, 128, 4), dtype=torch.uint8, device="cuda")
    ctx().rescale_bgra(dev(src)[None], out, wait=True)
    assert np.array_equal(out[0].cpu().numpy(), src)


@pytest.mark.parametrize("iw,ih,ow,oh,frames", [(100, 70, 150, 101, 3), (96, 130, 97, 131, 2), (33, 65, 1030, 67, 2)])
def test_rescale_batch_padded_pitches(iw, ih, ow, oh, frames):
    """several frames, row pitches and frame strides wider than the image, guard bytes untouched;
    near-unity scales make some source rows emit no output row (the walker must skip them)"""
    from _gpu import ctx, dev
    srcs = [bgra_random(iw, ih, seed=40 + f, alpha=None) for f in range(frames)]
    big_in = torch.full((frames, ih + 3, iw + 5, 4), 0x5A, dtype=torch.uint8, device="cuda")
    for f in range(frames):
        big_in[f, :ih, :iw] = dev(srcs[f])
    big_out = torch.full((frames, oh + 2, ow + 7, 4), 0xA5, dtype=torch.uint8, device="cuda")
    ctx().rescale_bgra(big_in[:, :ih, :iw], big_out[:, :oh, :ow], wait=True)
    got = big_out.cpu().numpy()
    for f in range(frames):
        _exact(got[f, :oh, :ow], ORA.rescale_bgra(srcs[f], ow, oh), f"frame {f}")
    assert (got[:, oh:, :, :] == 0xA5).all() and (got[:, :, ow:, :] == 0xA5).all(), "wrote outside the image"


def test_rescale_alpha_channel_is_linear():
    """A is blended without the sRGB curve (the texture's alpha is linear): constant alpha stays constant,
    and a vertical alpha ramp upscaled 2x matches the oracle"""
    from _gpu import ctx, dev
    src = bgra_random(64, 32, seed=9, alpha=None)
    src[..., 3] = np.arange(32, dtype=np.uint8)[:, None] * 8
    out = torch.zeros((1, 64, 128, 4), dtype=torch.uint8, device="cuda")
    ctx().rescale_bgra(dev(src)[None], out, wait=True)
    _exact(out[0].cpu().numpy(), ORA.rescale_bgra(src, 128, 64), "alpha ramp")
    src[..., 3] = 200
    ctx().rescale_bgra(dev(src)[None], out, wait=True)
    assert (out[0, :, :, 3].cpu().numpy() == 200).all()


def test_decode_rescale_policy():
    """O1 (AAPLRenderer.m:891-977): same size -> one pass; else decode -> intermediate -> rescale"""
    from _gpu import ctx, dev
    from metalbt709decoder_b200 import BT709Error
    C = ctx()
    w, h, ow, oh = 320, 180, 427, 240
    y, c = frame_random(w, h, seed=8)
    ty, tc = dev(y)[None], dev(c)[None]
    mid_want = ORA.decode_nv12(1, y, c)
    out = torch.zeros((1, h, w, 4), dtype=torch.uint8, device="cuda")
    C.decode_rescale_nv12(1, ty, tc, None, out, w=w, h=h, wait=True)  # same size: scratch may be NULL
    _exact(out[0].cpu().numpy(), mid_want, "one pass")
    out2 = torch.zeros((1, oh, ow, 4), dtype=torch.uint8, device="cuda")
    with pytest.raises(BT709Error) as e:
        C.decode_rescale_nv12(1, ty, tc, None, out2, w=w, h=h)
    assert e.value.status == -7
    scratch = torch.zeros((1, h, w, 4), dtype=torch.uint8, device="cuda")
    n0 = C.launch_count
    C.decode_rescale_nv12(1, ty, tc, scratch, out2, w=w, h=h, wait=True)
    # 320 is not a multiple of 128: two kernels through the same-size intermediate, which `scratch` then holds
    assert C.launch_count == n0 + 2
    _exact(scratch[0].cpu().numpy(), mid_want, "intermediate")
    _exact(out2[0].cpu().numpy(), ORA.rescale_bgra(mid_want, ow, oh), "two pass")


# ------------------------------------------------------------------------------ encode
PAIRS = [(GAMMA_SRGB, GAMMA_APPLE), (GAMMA_SRGB, GAMMA_SRGB), (GAMMA_LINEAR, GAMMA_LINEAR),
         (GAMMA_APPLE, GAMMA_SRGB), (GAMMA_SRGB, GAMMA_LINEAR), (GAMMA_APPLE, GAMMA_APPLE)]


def _encode(gin, gout, bgra, pad=0):
    from _gpu import ctx, dev
    h, w, _ = bgra.shape
    y = torch.zeros((1, h, w + pad), dtype=torch.uint8, device="cuda")
    c = torch.zeros((1, h // 2, w + pad), dtype=torch.uint8, device="cuda")
    ctx().encode_bgra_to_nv12(gin, gout, dev(bgra)[None], y, c, wait=True)
    return y[0, :, :w].cpu().numpy(), c[0, :, :w].cpu().numpy()


@pytest.mark.parametrize("gin,gout", PAIRS)
@pytest.mark.parametrize("w,h", [(2048, 64), (6, 4), (130, 2), (720, 6), (1440, 4), (132, 6), (4092, 2), (4, 2)])
def test_encode420_matches_oracle(gin, gout, w, h):
    bgra = bgra_random(w, h, seed=711)
    gy, gc = _encode(gin, gout, bgra, pad=(2 if w % 4 else 0))
    wy, wc = ORA.encode_bgra_nv12(gin, gout, bgra)
    _exact(gy, wy, "Y")
    _exact(gc, wc, "CbCr")


@pytest.mark.parametrize("w,h", [(854, 480), (1366, 8), (130, 4), (6, 2)])
def test_encode_even_widths_not_multiple_of_4_on_the_fast_path(w, h):
    """w % 4 == 2 with 16-byte-aligned BGRA pitch and 4-byte-aligned Y / CbCr pitches: the last lane of a row
    owns one 2x2 block and stores 2 bytes per plane row; bytes past the row end stay untouched"""
    from _gpu import ctx, dev
    F = 2
    imgs = [bgra_random(w, h, seed=800 + f) for f in range(F)]
    pw = w + 2
    tin = torch.full((F, h, pw, 4), 0x33, dtype=torch.uint8, device="cuda")
    for f in range(F):
        tin[f, :, :w] = dev(imgs[f])
    ty = torch.full((F, h, pw), 0xEE, dtype=torch.uint8, device="cuda")
    tc = torch.full((F, h // 2, pw), 0xDD, dtype=torch.uint8, device="cuda")
    ctx().encode_bgra_to_nv12(GAMMA_SRGB, GAMMA_APPLE, tin[:, :, :w], ty[:, :, :w], tc[:, :, :w], wait=True)
    gy, gc = ty.cpu().numpy(), tc.cpu().numpy()
    for f in range(F):
        wy, wc = ORA.encode_bgra_nv12(GAMMA_SRGB, GAMMA_APPLE, imgs[f])
        _exact(gy[f, :, :w], wy, f"Y frame {f}")
        _exact(gc[f, :, :w], wc, f"CbCr frame {f}")
    assert (gy[:, :, w:] == 0xEE).all() and (gc[:, :, w:] == 0xDD).all(), "wrote past the end of a row"


def test_encode_ramp_and_alpha_ignored():
    """F-ramp (gamma_write.m:280-298); the A byte is ignored (CVPixelBufferUtils.h:304-318)"""
    bgra = frame_ramp(1024, 8)
    a = bgra.copy(); a[..., 3] = 7
    y1, c1 = _encode(GAMMA_SRGB, GAMMA_APPLE, bgra)
    y2, c2 = _encode(GAMMA_SRGB, GAMMA_APPLE, a)
    wy, wc = ORA.encode_bgra_nv12(GAMMA_SRGB, GAMMA_APPLE, bgra)
    assert np.array_equal(y1, wy) and np.array_equal(c1, wc)
    assert np.array_equal(y1, y2) and np.array_equal(c1, c2)
    assert (c1 == 128).all()  # gray has neutral chroma


def test_encode_error_codes():
    from _gpu import ctx, dev
    from metalbt709decoder_b200 import BT709Error
    C = ctx()
    bgra = dev(bgra_random(6, 4))[None]
    y = torch.zeros((1, 4, 8), dtype=torch.uint8, device="cuda")
    c = torch.zeros((1, 2, 8), dtype=torch.uint8, device="cuda")
    with pytest.raises(BT709Error) as e:
        C.encode_bgra_to_nv12(3, 0, bgra, y, c)  # BT.709 spec curve is not defined for this path
    assert e.value.status == -4
    odd = dev(bgra_random(5, 4))[None]
    with pytest.raises(BT709Error) as e:
        C.encode_bgra_to_nv12(1, 0, odd, y, c)
    assert e.value.status == -2


@pytest.mark.parametrize("gin,gout,dec", [(GAMMA_SRGB, GAMMA_APPLE, 0), (GAMMA_SRGB, GAMMA_SRGB, 1),
                                          (GAMMA_LINEAR, GAMMA_LINEAR, 2)])
def test_encode_decode_roundtrip_histogram(gin, gout, dec):
    """C5: 2048x1536 encode then decode with the matching gamma.  The GPU round trip must be
    identical to the oracle's; the max-channel abs-error histogram against the source is reported."""
    from _gpu import decode
    w, h = 2048, 1536
    bgra = bgra_random(w, h, seed=711)
    gy, gc = _encode(gin, gout, bgra)
    got = decode(dec, gy, gc)
    wy, wc = ORA.encode_bgra_nv12(gin, gout, bgra)
    assert np.array_equal(gy, wy) and np.array_equal(gc, wc)
    want = ORA.decode_nv12(dec, wy, wc)
    assert np.array_equal(got, want)
    # flat 2x2 blocks remove the chroma-subsampling loss: what is left is quantisation only
    blocks = np.random.default_rng(711).integers(0, 256, (h // 2, w // 2, 4), dtype=np.uint8)
    flat = np.repeat(np.repeat(blocks, 2, axis=0), 2, axis=1)
    fy, fc = _encode(gin, gout, flat)
    back = decode(dec, fy, fc)
    d = np.abs(back[..., :3].astype(np.int16) - flat[..., :3].astype(np.int16)).max(axis=-1)
    hist = np.bincount(d.ravel(), minlength=8)
    print(f"round trip gin={gin} gout={gout}: max-channel |d| histogram on flat blocks:", hist[:8].tolist())
    if dec != 2:  # Linear data goes through the sRGB store conversion: not a round trip
        # CoreImageMetalFilterTests.m:539-553, 676-691: the 4:4:4 round trips stay within 5 LSB
        assert d.max() <= 5, hist.tolist()


# ---------------------------------------------------------------------- packed 4:4:4 harness
def test_packed444_convert_unconvert_copy():
    from _gpu import ctx, dev
    C = ctx()
    w, h = 64, 32
    words = (bgra_random(w, h, seed=9).view(np.uint32).reshape(h, w) & 0x00FFFFFF)
    rc, want = ORA.convert_packed444(words)
    assert rc == 0
    out = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    C.convert_packed444(dev(words.view(np.int32)), out, wait=True)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want)
    rc, back = ORA.unconvert_packed444(want)
    out2 = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    C.unconvert_packed444(out, out2, wait=True)
    assert np.array_equal(out2.cpu().numpy().view(np.uint32), back)
    wy, wc = ORA.copy_packed444_to_nv12(want)
    ty = torch.zeros((1, h, w), dtype=torch.uint8, device="cuda")
    tc = torch.zeros((1, h // 2, w), dtype=torch.uint8, device="cuda")
    C.copy_packed444_to_nv12(out, ty, tc, wait=True)
    assert np.array_equal(ty[0].cpu().numpy(), wy) and np.array_equal(tc[0].cpu().numpy(), wc)
