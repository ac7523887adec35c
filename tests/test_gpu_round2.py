"""Round-2 parity tests (-m gpu), all through the C ABI:
  * the GPU against oracle/_ref (the reference's own BT709.h compiled where it lies) DIRECTLY, so the parity chain
    GPU == reference closes inside the driver's -m gpu run and not only through the restatement;
  * the CoreImageMetalFilterTests.m goldens (five 4:2:0 averaging vectors, nine BT.709-curve pixel tests) replayed on
    the device, and the three exhaustive 2^24-colour round-trip histograms reproduced on the device;
  * the geometries of the linear-tile-stream decode kernel (k_decode_lin) and its few-replica variant;
  * the encode launcher's 32-bit offset split, and slot reuse of the host layer across caller streams.
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from _oracle import GAMMA_APPLE, GAMMA_BT709, GAMMA_LINEAR, GAMMA_SRGB, Oracle, bgra_random, frame_lattice, frame_random, ref_oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
VEC = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
ORA = Oracle()
GAMMA_BT709_LINEAR_RGB = 4


def _ref():
    r = ref_oracle()
    if r is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return r


# ------------------------------------------------------------------ GPU == reference, directly
@pytest.mark.parametrize("gamma", [0, 1, 2, 3])
def test_lattice_decode_equals_the_reference_headers(gamma):
    """all 2^24 (Y,Cb,Cr) inputs: bt709_decode_nv12 against BT709.h/sRGB.h compiled unmodified (oracle/_ref)"""
    from _gpu import decode
    ref = _ref()
    y, c = frame_lattice()
    got = decode(gamma, y, c)
    want = ref.decode_nv12(gamma, y, c)
    assert np.array_equal(got[..., :3], want[..., :3]), f"gamma {gamma}: {(got[..., :3] != want[..., :3]).sum()} mismatches"
    assert (got[..., 3] == 255).all()


@pytest.mark.parametrize("gin,gout", [(1, 0), (1, 1), (2, 2), (0, 0), (1, 2), (0, 1)])
def test_encode_equals_the_reference_headers(gin, gout):
    """bt709_encode_bgra_to_nv12 against BT709_average_pixel_values of the reference (oracle/_ref), six gamma pairs"""
    from _gpu import ctx, dev
    ref = _ref()
    w, h = 256, 64
    src = bgra_random(w, h, seed=900 + gin * 3 + gout)
    ty = torch.zeros((1, h, w), dtype=torch.uint8, device="cuda")
    tc = torch.zeros((1, h // 2, w), dtype=torch.uint8, device="cuda")
    ctx().encode_bgra_to_nv12(gin, gout, dev(src)[None], ty, tc, wait=True)
    wy, wc = ref.encode_bgra_nv12(gin, gout, src)
    assert np.array_equal(ty[0].cpu().numpy(), wy) and np.array_equal(tc[0].cpu().numpy(), wc)


# ------------------------------------------------------------------ CoreImageMetalFilterTests.m on the device
CI = {t["name"]: t for t in VEC["files"]["CoreImageMetalFilterTests.m"]}
REF_GAMMA = {"BT709GammaSrgb": GAMMA_SRGB, "BT709GammaApple": GAMMA_APPLE, "BT709GammaLinear": GAMMA_LINEAR}


def _chk(t):
    return {c["var"]: c["expected"] for c in t["checks"]}


def _convert444(gamma, r, g, b):
    """one colour as a 2x2 frame through bt709_convert_packed444_gamma -> (Y, Cb, Cr)"""
    from _gpu import ctx
    words = torch.full((2, 2), (0xFF << 24) | (r << 16) | (g << 8) | b, dtype=torch.int64).to(torch.int32).cuda()
    out = torch.zeros((2, 2), dtype=torch.int32, device="cuda")
    ctx().convert_packed444(words, out, wait=True, gamma=gamma)
    v = out.cpu().numpy().astype(np.uint32)
    assert (v == v[0, 0]).all()
    p = int(v[0, 0])
    return p & 255, (p >> 8) & 255, (p >> 16) & 255


def _unconvert444(gamma, y, cb, cr):
    from _gpu import ctx
    words = torch.full((2, 2), (cr << 16) | (cb << 8) | y, dtype=torch.int32, device="cuda")
    out = torch.zeros((2, 2), dtype=torch.int32, device="cuda")
    ctx().unconvert_packed444(words, out, wait=True, gamma=gamma)
    p = int(out.cpu().numpy().astype(np.uint32)[1, 1])
    return (p >> 16) & 255, (p >> 8) & 255, p & 255


def _decode_2x2(gamma, y, cb, cr):
    from _gpu import decode
    got = decode(gamma, np.full((2, 2), y, np.uint8), np.array([[cb, cr]], np.uint8))
    assert (got == got[0, 0]).all()
    return int(got[0, 0, 2]), int(got[0, 0, 1]), int(got[0, 0, 0])


@pytest.mark.parametrize("name", [n for n in CI if n.startswith("testConvertsLinearRGBToBT709_") and "srgb" not in n])
def test_linear_rgb_to_bt709_goldens_on_the_device(name):
    """CoreImageMetalFilterTests.m:94-360: BT709_convertRGBToYCbCr / BT709_convertYCbCrToRGB with and without the
    BT.709 curve.  Without it the pair is byteNorm -> matrix -> round, i.e. the sRGB-tagged pair."""
    t = CI[name]
    c, i = _chk(t), t["inputs"]
    g = GAMMA_BT709_LINEAR_RGB if i["applyGammaMap"] else GAMMA_SRGB
    assert _convert444(g, i["Rin"], i["Gin"], i["Bin"]) == (c["Y"], c["Cb"], c["Cr"])
    assert _unconvert444(g, c["Y"], c["Cb"], c["Cr"]) == (c["R"], c["G"], c["B"])


@pytest.mark.parametrize("name", ["testConvertsLinearRGBToBT709_42PercentBlue_srgb",
                                  "testConvertsSRGBToYCbCr_ResampleUpperNW", "testConvertsSRGBToYCbCr_ResampleUpperNE",
                                  "testConvertsSRGBToYCbCr_ResampleUpperSW", "testConvertsSRGBToYCbCr_ResampleUpperSE"])
def test_bt709_spec_curve_goldens_on_the_device(name):
    """CoreImageMetalFilterTests.m:362-416, 851-1073: BT709_from_sRGB_convertRGBToYCbCr on the device (E1 + E4), then
    BT709_to_sRGB_convertYCbCrToRGB both through the packed harness kernel and through bt709_decode_nv12"""
    t = CI[name]
    c, i = _chk(t), t["inputs"]
    assert _convert444(GAMMA_BT709, i["Rin"], i["Gin"], i["Bin"]) == (c["Y"], c["Cb"], c["Cr"])
    assert _unconvert444(GAMMA_BT709, c["Y"], c["Cb"], c["Cr"]) == (c["R"], c["G"], c["B"])
    assert _decode_2x2(GAMMA_BT709, c["Y"], c["Cb"], c["Cr"]) == (c["R"], c["G"], c["B"])


@pytest.mark.parametrize("name", [n for n in CI if n.startswith("testConvertsSRGBToYCbCr_AverageOf4")])
def test_average_of_4_goldens_on_the_device(name):
    """CoreImageMetalFilterTests.m:1683-2094: BT709_average_pixel_values vectors as one 2x2 frame through
    bt709_encode_bgra_to_nv12"""
    from _gpu import ctx
    t = CI[name]
    c, i = _chk(t), t["inputs"]
    gin, gout = REF_GAMMA[t["gammas"]["inputGamma"]], REF_GAMMA[t["gammas"]["outputGamma"]]
    px = np.zeros((2, 2, 4), np.uint8)
    for k, (row, col) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)], start=1):  # NW, NE, SW, SE
        px[row, col] = (i[f"B{k}"], i[f"G{k}"], i[f"R{k}"], 255)
    ty = torch.zeros((1, 2, 2), dtype=torch.uint8, device="cuda")
    tc = torch.zeros((1, 1, 2), dtype=torch.uint8, device="cuda")
    ctx().encode_bgra_to_nv12(gin, gout, torch.from_numpy(px).cuda()[None], ty, tc, wait=True)
    assert ty[0].cpu().numpy().reshape(-1).tolist() == [c["Y1"], c["Y2"], c["Y3"], c["Y4"]]
    assert tc[0, 0].cpu().numpy().tolist() == [c["Cb"], c["Cr"]]


@pytest.mark.parametrize("key", ["bt709", "apple", "srgb"])
def test_2_24_roundtrip_histograms_on_the_device(key):
    """CoreImageMetalFilterTests.m:420-537, 557-674, 696-813: every sRGB colour through the per-pixel encoder and
    decoder of one curve, bucketed by the largest channel difference; every bucket must match the reference's
    recorded counts (all three encoders and decoders run on the GPU)"""
    from _gpu import ctx
    hspec = VEC["histograms"][key]
    g = hspec["gamma"]
    n = torch.arange(1 << 24, dtype=torch.int32, device="cuda").view(4096, 4096)
    ycc = torch.empty_like(n)
    back = torch.empty_like(n)
    ctx().convert_packed444(n, ycc, gamma=g)
    ctx().unconvert_packed444(ycc, back, gamma=g, wait=True)
    dmax = torch.zeros_like(n)
    for sh in (0, 8, 16):
        d = (((n >> sh) & 255) - ((back >> sh) & 255)).abs()
        dmax = torch.maximum(dmax, d)
    hist = torch.bincount(dmax.view(-1), minlength=11).cpu().tolist()
    assert hist[0] == hspec["exact"]
    assert hist[:len(hspec["buckets"])] == hspec["buckets"]
    assert sum(hist) == 1 << 24 and sum(hist[len(hspec["buckets"]):]) == 0


@pytest.mark.parametrize("gamma", [GAMMA_APPLE, GAMMA_SRGB, GAMMA_BT709, GAMMA_BT709_LINEAR_RGB])
def test_packed444_gamma_variants_match_the_oracle(gamma):
    from _gpu import ctx
    rng = np.random.default_rng(1234 + gamma)
    h, w = 32, 48
    words = rng.integers(0, 1 << 24, (h, w), dtype=np.uint32)
    tin = torch.from_numpy(words.astype(np.int32)).cuda()
    tout = torch.zeros_like(tin)
    ctx().convert_packed444(tin, tout, wait=True, gamma=gamma)
    got = tout.cpu().numpy().astype(np.uint32)
    back = torch.zeros_like(tin)
    ctx().unconvert_packed444(tout, back, wait=True, gamma=gamma)
    gotb = back.cpu().numpy().astype(np.uint32)
    for (r_, c_) in [(int(a), int(b)) for a, b in zip(rng.integers(0, h, 300), rng.integers(0, w, 300))]:
        p = int(words[r_, c_])
        R, G, B = (p >> 16) & 255, (p >> 8) & 255, p & 255
        if gamma == GAMMA_BT709_LINEAR_RGB:
            Y, Cb, Cr = ORA.linear_rgb_to_ycbcr(R, G, B, 1)
            rr = ORA.ycbcr_to_linear_rgb(Y, Cb, Cr, 1)
        else:
            Y, Cb, Cr = ORA.encode444_pixel(gamma, R, G, B)
            rr = ORA.decode_pixel(gamma, Y, Cb, Cr)
        assert int(got[r_, c_]) == (Cr << 16) | (Cb << 8) | Y
        q = int(gotb[r_, c_])
        assert ((q >> 16) & 255, (q >> 8) & 255, q & 255) == rr and (q >> 24) == 0


def test_packed444_gamma_errors():
    from _gpu import ctx
    from metalbt709decoder_b200 import BT709Error
    a = torch.zeros((2, 2), dtype=torch.int32, device="cuda")
    for g in (2, 5, -1):
        with pytest.raises(BT709Error) as e:
            ctx().convert_packed444(a, a.clone(), gamma=g)
        assert e.value.status == -4
        with pytest.raises(BT709Error) as e:
            ctx().unconvert_packed444(a, a.clone(), gamma=g)
        assert e.value.status == -4


# ------------------------------------------------------------------ k_decode_lin geometries
LIN_CASES = [
    # w, h, frames, gamma, alpha
    (1920, 1080, 3, 3, False),   # 15 tiles per row, replicated LUT (24300 tiles)
    (1920, 1080, 3, 2, False),   # Linear mode: log-cell LUT, compact ring slices, 30 warps per CTA
    (2048, 1536, 2, 1, True),    # alpha stream, 16 tiles per row
    (1920, 1080, 1, 0, False),   # one frame per call: the few-replica variant
    (1920, 2, 1, 3, False),      # a single row pair
    (128, 2, 1, 1, False),       # a single tile
    (128, 64, 40, 0, False),     # one tile per row, many frames
    (4224, 4, 2, 3, False),      # 33 tiles per row (more than the warps of a CTA)
    (8192, 8, 2, 2, False),      # 64 tiles per row, Linear mode
    (3840, 6, 5, 1, True),       # UHD width with alpha, small call
    (640, 360, 9, 3, False),     # 5 tiles per row
    # ragged widths: a partial last tile column (per-lane copy sizes, 16- / 8- / 0-byte stores)
    (720, 480, 3, 3, False),     # 5 full tiles + 80 pixels
    (960, 540, 2, 2, False),     # 7 full tiles + 64 pixels, Linear mode
    (1440, 1080, 1, 0, False),   # 11 full tiles + 32 pixels, one frame per call (few-replica LUT)
    (352, 288, 4, 1, True),      # 2 full tiles + 96 pixels, alpha stream
    (48, 4, 3, 3, False),        # less than one tile
    (4240, 4, 2, 1, False),      # 33 full tiles + 16 pixels
]


@pytest.mark.parametrize("w,h,F,gamma,alpha", LIN_CASES)
def test_linear_tile_stream_kernel(w, h, F, gamma, alpha):
    from _gpu import ctx, dev
    frames = [frame_random(w, h, seed=500 + f, full_range=(f % 2 == 1)) for f in range(F)]
    ty = torch.stack([dev(y) for y, _ in frames])
    tc = torch.stack([dev(c) for _, c in frames])
    alphas, ta = None, None
    if alpha:
        alphas = [frame_random(w, h, seed=600 + f, full_range=True)[0] for f in range(F)]
        ta = torch.stack([dev(a) for a in alphas])
    out = torch.zeros((F, h, w, 4), dtype=torch.uint8, device="cuda")
    n0 = ctx().launch_count
    ctx().decode_nv12(gamma, ty, tc, out, alpha_y=ta, wait=True)
    assert ctx().launch_count == n0 + 1
    got = out.cpu().numpy()
    for f in range(F):
        want = ORA.decode_nv12(gamma, frames[f][0], frames[f][1], alpha_y=alphas[f] if alpha else None)
        assert np.array_equal(got[f], want), f"{w}x{h} frame {f}: {(got[f] != want).sum()} byte mismatches"


@pytest.mark.parametrize("env", [{"BT709_SMALL_STEPS": "0"}, {"BT709_SMALL_STEPS": "1000000"}, {"BT709_NO_LIN2": "1"},
                                 {"BT709_RING": "3l"}, {"BT709_RING": "2e"}, {"BT709_FUSED": "1"}])
def test_decode_lut_replication_variants_in_a_subprocess(env):
    """BT709_SMALL_STEPS=0: the per-lane replicated LUT even for tiny calls; =1000000: the few-replica LUT even for
    batches; BT709_NO_LIN2: the round-1 kernels; BT709_RING: both ring variants of k_decode_lin for every mode;
    BT709_FUSED: the policy entry takes the fused kernel.  Each must stay exact (env is read once per process)."""
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "tests")
from _oracle import Oracle, frame_random
from _gpu import ctx, dev
ora = Oracle()
for (w, h, F, g, al) in [(1920, 64, 2, 3, False), (3840, 8, 3, 2, False), (2048, 16, 2, 1, True), (384, 6, 4, 0, False), (4224, 4, 2, 2, False)]:
    fr = [frame_random(w, h, seed=70 + f) for f in range(F)]
    ty = torch.stack([dev(y) for y, _ in fr]); tc = torch.stack([dev(c) for _, c in fr])
    al_np = [frame_random(w, h, seed=80 + f, full_range=True)[0] for f in range(F)] if al else None
    ta = torch.stack([dev(a) for a in al_np]) if al else None
    out = torch.zeros((F, h, w, 4), dtype=torch.uint8, device="cuda")
    ctx().decode_nv12(g, ty, tc, out, alpha_y=ta, wait=True)
    got = out.cpu().numpy()
    for f in range(F):
        assert np.array_equal(got[f], ora.decode_nv12(g, fr[f][0], fr[f][1], alpha_y=al_np[f] if al else None)), (w, h, g, f)
from _oracle import bgra_random
w, h, ow, oh = 256, 64, 400, 100
y, c = frame_random(w, h, seed=91)
out = torch.zeros((1, oh, ow, 4), dtype=torch.uint8, device="cuda")
ctx().decode_rescale_nv12(3, dev(y)[None], dev(c)[None], torch.zeros((1, h, w, 4), dtype=torch.uint8, device="cuda"), out, wait=True)
assert np.array_equal(out[0].cpu().numpy(), ora.rescale_bgra(ora.decode_nv12(3, y, c), ow, oh))
print("variant ok")
'''
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, **env), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout + r.stderr


# ------------------------------------------------------------------ encode: offsets beyond 32-bit units
def test_encode_output_offsets_beyond_32_bits_are_split():
    """Y / CbCr planes with a padded frame stride: the batch's Y offsets pass 2^32 four-byte units (16 GiB) although
    the BGRA input is tiny.  The fast encode kernel addresses every plane with 32-bit unit offsets, so the launcher
    must split the batch; before the fix frames past the wrap were written into the wrong place."""
    from _gpu import ctx
    from metalbt709decoder_b200._lib import Plane
    w, h, F = 256, 8, 5
    stride_y = (9 << 29) + 4096      # 4.5 GiB between Y frames: frame 4 starts at 18 GiB
    stride_c = (5 << 30) + 8192      # 5 GiB between CbCr frames
    free, _ = torch.cuda.mem_get_info()
    need = (stride_y + stride_c) * (F - 1) + (1 << 30)
    if free < need:
        pytest.skip("not enough free device memory")
    ybuf = torch.empty(stride_y * (F - 1) + w * h, dtype=torch.uint8, device="cuda")
    cbuf = torch.empty(stride_c * (F - 1) + w * h // 2, dtype=torch.uint8, device="cuda")
    src = np.stack([bgra_random(w, h, seed=820 + f) for f in range(F)])
    tsrc = torch.from_numpy(src).cuda()
    for f in range(F):
        ybuf[f * stride_y:f * stride_y + w * h] = 0
        cbuf[f * stride_c:f * stride_c + w * h // 2] = 0
    C = ctx()
    pi = Plane(tsrc.data_ptr(), w * 4, w * h * 4)
    py = Plane(ybuf.data_ptr(), w, stride_y)
    pc = Plane(cbuf.data_ptr(), w, stride_c)
    n0 = C.launch_count
    rc = C._L.bt709_encode_bgra_to_nv12(C.handle, 1, 0, ctypes.byref(pi), ctypes.byref(py), ctypes.byref(pc), w, h, F,
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), 1)
    assert rc == 0
    assert C.launch_count - n0 >= 2, "the batch must have been split across launches"
    for f in range(F):
        wy, wc = ORA.encode_bgra_nv12(1, 0, src[f])
        gy = ybuf[f * stride_y:f * stride_y + w * h].view(h, w).cpu().numpy()
        gc = cbuf[f * stride_c:f * stride_c + w * h // 2].view(h // 2, w).cpu().numpy()
        assert np.array_equal(gy, wy) and np.array_equal(gc, wc), f"frame {f}"


# ------------------------------------------------------------------ host layer: slot reuse across streams
def test_host_layer_mixes_caller_streams_and_internal_streams():
    """More calls than staging slots on one MetalRenderContext, alternating a caller stream (commandBuffer) and NULL:
    a slot's reuse must always be ordered after its previous work (it was only ordered when the NEW call
    brought its own stream)"""
    from metalbt709decoder_b200 import renderer as R
    mrc = R.MetalRenderContext(0)
    assert mrc.setupMetal()
    dec = R.MetalBT709Decoder()
    dec.metalRenderContext = mrc
    dec.gamma = R.MetalBT709GammaSRGB
    w, h, n = 1920, 1080, 13
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    bufs, texs, wants = [], [], []
    for i in range(n):
        y, c = frame_random(w, h, seed=1000 + i)
        pb = R.BGRAToBT709Converter.createCoreVideoYCbCrBuffer(w, h)
        pb.y[:] = y
        pb.cbcr[:] = c
        pb.setAttachments(R.kMBDYCbCrMatrix_ITU_R_709_2, R.kMBDTransferFunction_sRGB)
        bufs.append(pb)
        texs.append(mrc.makeBGRATexture(w, h))
        wants.append((y, c))
    for i in range(n):
        cb = None if i % 3 == 0 else streams[i % 2].cuda_stream
        assert dec.decodeBT709(bufs[i], None, texs[i], cb, w, h, False)
    for s in streams:
        s.synchronize()
    assert mrc.waitUntilCompleted()
    for i in (0, 1, 4, 5, 8, 12):
        got = mrc.getBGRATexturePixels(texs[i]).view(np.uint8).reshape(h, w, 4)
        assert np.array_equal(got, ORA.decode_nv12(1, *wants[i])), f"call {i}"
    for b in bufs:
        b.release()
    for t in texs:
        t.release()
    dec.release()
    mrc.release()


# ------------------------------------------------------------------ pass 2: paired-column kernel (k_rescale_pair)
@pytest.mark.parametrize("mode", ["opaque", "mixed", "alpha"])
@pytest.mark.parametrize("iw,ih,ow,oh,F", [(192, 108, 256, 144, 3), (300, 200, 300, 200, 2), (64, 48, 200, 150, 2),
                                          (130, 70, 131, 141, 2), (640, 360, 1280, 720, 2), (96, 500, 97, 650, 1)])
def test_rescale_paired_columns_opaque_speculation(mode, iw, ih, ow, oh, F):
    """k_rescale_pair drops the alpha channel while every tap it sees has A = 255 and redoes a tile with the
    four-channel code as soon as one does not; the next tiles of that warp then start on the alpha code and fall
    back to the opaque code on the first fully opaque tile.  Opaque, fully translucent and mixed inputs (a few
    translucent patches, so that both switches happen inside one launch) must all match the oracle."""
    from _gpu import ctx, dev
    rng = np.random.default_rng(77)
    srcs = []
    for f in range(F):
        a = 255 if mode in ("opaque", "mixed") else None
        src = bgra_random(iw, ih, seed=300 + f, alpha=a)
        if mode == "mixed":
            for _ in range(3):
                x0, y0 = int(rng.integers(0, iw)), int(rng.integers(0, ih))
                src[y0:y0 + 7, x0:x0 + 5, 3] = rng.integers(0, 255, src[y0:y0 + 7, x0:x0 + 5, 3].shape, dtype=np.uint8)
        srcs.append(src)
    tin = torch.stack([dev(s_) for s_ in srcs])
    out = torch.zeros((F, oh, ow, 4), dtype=torch.uint8, device="cuda")
    ctx().rescale_bgra(tin, out, wait=True)
    got = out.cpu().numpy()
    for f in range(F):
        want = ORA.rescale_bgra(srcs[f], ow, oh)
        assert np.array_equal(got[f], want), f"{mode} {iw}x{ih}->{ow}x{oh} frame {f}: {(got[f] != want).sum()} byte mismatches"
        if mode == "opaque":
            assert (got[f][..., 3] == 255).all()


def test_rescale_1080p_to_1440p_opaque_full_size():
    """BASELINE configs[1] pass 2 at full size on what pass 1 hands over (A = 255)"""
    from _gpu import ctx, dev
    src = bgra_random(1920, 1080, seed=31, alpha=255)
    out = torch.zeros((1, 1440, 2560, 4), dtype=torch.uint8, device="cuda")
    ctx().rescale_bgra(dev(src)[None], out, wait=True)
    assert np.array_equal(out[0].cpu().numpy(), ORA.rescale_bgra(src, 2560, 1440))


# ------------------------------------------------------------------ N1: fused decode -> rescale
FUSED_CASES = [
    # w, h, ow, oh, frames, gamma
    (1920, 1080, 2560, 1440, 2, 3),   # BASELINE configs[1]
    (1920, 1080, 2560, 1440, 1, 2),   # Linear mode: the decode reads the encode LUT; one frame (fewer slices than SMs)
    (1280, 720, 1920, 1080, 3, 0),
    (640, 360, 640, 481, 2, 1),       # same width, vertical upscale only
    (256, 64, 258, 40, 4, 3),         # vertical downscale (1.6 source rows per output row)
    (128, 2, 300, 7, 3, 1),           # a single row pair
    (384, 130, 1000, 333, 2, 0),
    (1920, 1080, 3840, 2160, 1, 3),
]


@pytest.mark.parametrize("w,h,ow,oh,F,gamma", FUSED_CASES)
def test_fused_decode_rescale_equals_decode_then_rescale(w, h, ow, oh, F, gamma):
    """bt709_decode_rescale_nv12 as ONE kernel: bit-identical to the oracle's decode followed by the
    oracle's rescale, one launch, and the intermediate image is not in `scratch` afterwards"""
    from _gpu import ctx, dev
    frames = [frame_random(w, h, seed=1300 + f, full_range=(f % 2 == 1)) for f in range(F)]
    ty = torch.stack([dev(y) for y, _ in frames])
    tc = torch.stack([dev(c) for _, c in frames])
    scratch = torch.zeros((F, h, w, 4), dtype=torch.uint8, device="cuda")
    out = torch.zeros((F, oh, ow, 4), dtype=torch.uint8, device="cuda")
    C = ctx()
    n0 = C.launch_count
    if fused_fits_ := (F * h * w * 4 >= (int(64 * h / oh) + 12) * ((w * 4 + 255) // 256 * 256)):
        C.decode_rescale_nv12_fused(gamma, ty, tc, scratch, out, wait=True)
    else:
        C.decode_rescale_nv12(gamma, ty, tc, scratch, out, wait=True)
    # one launch = the fused kernel.  (A scratch area smaller than one CTA slice - the 128x2 case: 3 KB - cannot host
    # it; the call then runs the two passes.)
    assert C.launch_count == n0 + (1 if fused_fits_ else 2), "expected the fused kernel (one launch)"
    got = out.cpu().numpy()
    for f in range(F):
        mid = ORA.decode_nv12(gamma, frames[f][0], frames[f][1])
        want = ORA.rescale_bgra(mid, ow, oh)
        assert np.array_equal(got[f], want), f"frame {f}: {(got[f] != want).sum()} byte mismatches"


def test_fused_entry_refuses_what_it_cannot_do_and_the_policy_entry_runs_two_passes():
    """alpha stream / ragged width / padded scratch rows: the _fused entry returns BT709_E_UNSUPPORTED (-8), the policy
    entry bt709_decode_rescale_nv12 runs the two kernels (also with a full-size scratch: faster on B200)"""
    from _gpu import ctx, dev
    from metalbt709decoder_b200 import BT709Error
    C = ctx()
    w, h, ow, oh = 256, 64, 400, 100
    y, c = frame_random(w, h, seed=5)
    a, _ = frame_random(w, h, seed=6, full_range=True)
    ty, tc, ta = dev(y)[None], dev(c)[None], dev(a)[None]
    out = torch.zeros((1, oh, ow, 4), dtype=torch.uint8, device="cuda")
    scratch = torch.zeros((1, h, w, 4), dtype=torch.uint8, device="cuda")
    n0 = C.launch_count
    C.decode_rescale_nv12(1, ty, tc, scratch, out, alpha_y=ta, wait=True)
    assert C.launch_count == n0 + 2
    mid = ORA.decode_nv12(1, y, c, alpha_y=a)
    assert np.array_equal(scratch[0].cpu().numpy(), mid)
    assert np.array_equal(out[0].cpu().numpy(), ORA.rescale_bgra(mid, ow, oh))
    padded = torch.zeros((1, h, w + 8, 4), dtype=torch.uint8, device="cuda")[:, :, :w]
    with pytest.raises(BT709Error) as e:
        C.decode_rescale_nv12_fused(3, ty, tc, padded, out)
    assert e.value.status == -8
    yr, cr = frame_random(130, 64, seed=7)
    with pytest.raises(BT709Error) as e:
        C.decode_rescale_nv12_fused(3, dev(yr)[None], dev(cr)[None], torch.zeros((1, 64, 130, 4), dtype=torch.uint8, device="cuda"), out)
    assert e.value.status == -8
    # full-size scratch: the policy entry decodes into it and rescales from it (two launches)
    want = ORA.rescale_bgra(ORA.decode_nv12(3, y, c), ow, oh)
    n0 = C.launch_count
    C.decode_rescale_nv12(3, ty, tc, scratch, out, wait=True)
    assert C.launch_count == n0 + 2 and np.array_equal(out[0].cpu().numpy(), want)
    # two frames, scratch for one (frame stride 0): the fused entry works inside the first frame's area
    F = 2
    ty2, tc2 = ty.repeat(F, 1, 1), tc.repeat(F, 1, 1)
    out2 = torch.zeros((F, oh, ow, 4), dtype=torch.uint8, device="cuda")
    from metalbt709decoder_b200._lib import Plane
    ps = Plane(scratch.data_ptr(), w * 4, 0)
    py, pc, po = Plane(ty2.data_ptr(), w, w * h), Plane(tc2.data_ptr(), w, w * h // 2), Plane(out2.data_ptr(), ow * 4, ow * oh * 4)
    n0 = C.launch_count
    rc = C._L.bt709_decode_rescale_nv12_fused(C.handle, 3, ctypes.byref(py), ctypes.byref(pc), ctypes.byref(ps), ctypes.byref(po),
                                              w, h, ow, oh, F, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), 1)
    assert rc == 0 and C.launch_count == n0 + 1
    got = out2.cpu().numpy()
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)


def test_fused_full_batch_checksum_against_two_passes():
    """a 1080p -> 1440p batch larger than the number of SMs x bands per slice: the fused kernel against the two
    separate C-ABI calls on the same inputs (every byte)"""
    from _gpu import ctx
    C = ctx()
    F, w, h, ow, oh = 40, 1920, 1080, 2560, 1440
    g = torch.Generator(device="cuda")
    g.manual_seed(99)
    ty = torch.randint(16, 236, (F, h, w), dtype=torch.uint8, device="cuda", generator=g)
    tc = torch.randint(16, 241, (F, h // 2, w), dtype=torch.uint8, device="cuda", generator=g)
    mid = torch.empty((F, h, w, 4), dtype=torch.uint8, device="cuda")
    want = torch.empty((F, oh, ow, 4), dtype=torch.uint8, device="cuda")
    C.decode_nv12(3, ty, tc, mid)
    C.rescale_bgra(mid, want, wait=True)
    got = torch.zeros_like(want)
    scratch = torch.empty((F, h, w, 4), dtype=torch.uint8, device="cuda")
    n0 = C.launch_count
    C.decode_rescale_nv12_fused(3, ty, tc, scratch, got, wait=True)
    assert C.launch_count == n0 + 1
    assert torch.equal(got, want)


# ------------------------------------------------------------------ calls of a few frames (session 24)
@pytest.mark.parametrize("F", [1, 4, 9])
def test_rescale_tile_height_follows_the_batch_size(F):
    """k_rescale_pair halves its tile height (64 -> 32 -> 16 output rows) while a call has fewer tiles than the grid has
    warps, and deals the tiles out CTA by CTA: 1, 4 and 9 frames of 1080p -> 1440p run with 16-, 32- and 64-row tiles.
    Every frame must equal the oracle (two distinct sources, alternating)."""
    from _gpu import ctx, dev
    srcs = [bgra_random(1920, 1080, seed=41 + i, alpha=255) for i in range(2)]
    wants = [ORA.rescale_bgra(s_, 2560, 1440) for s_ in srcs]
    tin = torch.stack([dev(srcs[f % 2]) for f in range(F)])
    out = torch.zeros((F, 1440, 2560, 4), dtype=torch.uint8, device="cuda")
    ctx().rescale_bgra(tin, out, wait=True)
    for f in range(F):
        assert torch.equal(out[f].cpu(), torch.from_numpy(wants[f % 2])), f"frame {f} of {F}"


def test_back_to_back_small_calls_are_ordered():
    """Small calls are launched with programmatic stream serialization (the next kernel may start while the previous one
    drains) and wait in the kernel before they touch global memory.  A chain decode -> rescale -> encode per frame, twelve
    frames back to back on one stream with NO host synchronisation, the decode output and the rescaled image in buffers
    that every iteration overwrites (write-after-read against the previous iteration's consumer), must give the oracle's
    result for every frame."""
    from _gpu import ctx, dev
    w, h, ow, oh, n = 256, 64, 384, 96, 12
    c = ctx()
    frames = [frame_random(w, h, seed=900 + i) for i in range(n)]
    ty = [dev(y)[None] for y, _ in frames]
    tc = [dev(cc)[None] for _, cc in frames]
    mid = torch.zeros((1, h, w, 4), dtype=torch.uint8, device="cuda")     # shared by all iterations
    big = torch.zeros((1, oh, ow, 4), dtype=torch.uint8, device="cuda")   # shared by all iterations
    ey = [torch.zeros((1, oh, ow), dtype=torch.uint8, device="cuda") for _ in range(n)]
    ec = [torch.zeros((1, oh // 2, ow), dtype=torch.uint8, device="cuda") for _ in range(n)]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(n):
            c.decode_nv12(GAMMA_BT709, ty[i], tc[i], mid, stream=s)
            c.rescale_bgra(mid, big, stream=s)
            c.encode_bgra_to_nv12(GAMMA_SRGB, GAMMA_APPLE, big, ey[i], ec[i], stream=s)
    s.synchronize()
    for i in range(n):
        y, cc = frames[i]
        want_big = ORA.rescale_bgra(ORA.decode_nv12(GAMMA_BT709, y, cc), ow, oh)
        wy, wc = ORA.encode_bgra_nv12(GAMMA_SRGB, GAMMA_APPLE, want_big)
        assert np.array_equal(ey[i][0].cpu().numpy(), wy) and np.array_equal(ec[i][0].cpu().numpy(), wc), f"frame {i}"
    # the same chain through the two-pass policy entry and the fused entry
    with torch.cuda.stream(s):
        for i in range(n):
            c.decode_rescale_nv12(GAMMA_BT709, ty[i], tc[i], mid, big, stream=s)
            c.encode_bgra_to_nv12(GAMMA_SRGB, GAMMA_APPLE, big, ey[i], ec[i], stream=s)
    s.synchronize()
    for i in (0, 5, n - 1):
        y, cc = frames[i]
        wy, wc = ORA.encode_bgra_nv12(GAMMA_SRGB, GAMMA_APPLE, ORA.rescale_bgra(ORA.decode_nv12(GAMMA_BT709, y, cc), ow, oh))
        assert np.array_equal(ey[i][0].cpu().numpy(), wy) and np.array_equal(ec[i][0].cpu().numpy(), wc), f"two-pass frame {i}"


def test_small_calls_captured_in_a_cuda_graph():
    """The same chain (decode -> rescale -> encode per frame, shared intermediates) captured once into a CUDA graph and
    replayed: launches with programmatic stream serialization become programmatic dependency edges of the graph; the
    replay must still give the oracle's result for every frame, and again after the inputs have been changed in place."""
    from _gpu import ctx, dev
    w, h, ow, oh, n = 256, 64, 384, 96, 6
    c = ctx()
    frames = [frame_random(w, h, seed=950 + i) for i in range(n)]
    ty = [dev(y)[None] for y, _ in frames]
    tc = [dev(cc)[None] for _, cc in frames]
    mid = torch.zeros((1, h, w, 4), dtype=torch.uint8, device="cuda")
    big = torch.zeros((1, oh, ow, 4), dtype=torch.uint8, device="cuda")
    ey = [torch.zeros((1, oh, ow), dtype=torch.uint8, device="cuda") for _ in range(n)]
    ec = [torch.zeros((1, oh // 2, ow), dtype=torch.uint8, device="cuda") for _ in range(n)]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):  # warm-up outside the capture (module loading, attribute calls)
        c.decode_nv12(GAMMA_BT709, ty[0], tc[0], mid, stream=s)
        c.rescale_bgra(mid, big, stream=s)
        c.encode_bgra_to_nv12(GAMMA_SRGB, GAMMA_APPLE, big, ey[0], ec[0], stream=s)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        cur = torch.cuda.current_stream()
        for i in range(n):
            c.decode_nv12(GAMMA_BT709, ty[i], tc[i], mid, stream=cur)
            c.rescale_bgra(mid, big, stream=cur)
            c.encode_bgra_to_nv12(GAMMA_SRGB, GAMMA_APPLE, big, ey[i], ec[i], stream=cur)

    def check(fr, tag):
        for i in range(n):
            y, cc = fr[i]
            wy, wc = ORA.encode_bgra_nv12(GAMMA_SRGB, GAMMA_APPLE, ORA.rescale_bgra(ORA.decode_nv12(GAMMA_BT709, y, cc), ow, oh))
            assert np.array_equal(ey[i][0].cpu().numpy(), wy) and np.array_equal(ec[i][0].cpu().numpy(), wc), f"{tag} frame {i}"

    for t in ey + ec:
        t.zero_()
    g.replay()
    torch.cuda.synchronize()
    check(frames, "first replay")
    frames2 = [frame_random(w, h, seed=990 + i) for i in range(n)]
    for i in range(n):
        ty[i].copy_(dev(frames2[i][0])[None])
        tc[i].copy_(dev(frames2[i][1])[None])
    g.replay()
    torch.cuda.synchronize()
    check(frames2, "second replay")


@pytest.mark.parametrize("w,pitch,h,F,gamma,alpha", [(854, 896, 480, 2, 3, False), (1366, 1376, 768, 1, 0, False),
                                                    (854, 864, 6, 3, 1, True), (1366, 1376, 10, 2, 2, False)])
def test_ragged_widths_with_padded_pitches(w, pitch, h, F, gamma, alpha):
    """Even widths that are not a multiple of 16 in 16-byte-aligned pitched planes (what a hardware decoder hands over):
    the last 16-byte chunk of a row is copied with a source size of 6 or 14 bytes and the last lane stores 8 bytes
    (w % 4 == 2).  The padding of the output rows must stay untouched."""
    from _gpu import ctx, dev
    frames = [frame_random(w, h, seed=1200 + f, full_range=(f % 2 == 1)) for f in range(F)]
    ybuf = torch.zeros((F, h, pitch), dtype=torch.uint8, device="cuda")
    cbuf = torch.zeros((F, h // 2, pitch), dtype=torch.uint8, device="cuda")
    for f, (y, c) in enumerate(frames):
        ybuf[f, :, :w] = dev(y)
        cbuf[f, :, :w] = dev(c)
    alphas, abuf = None, None
    if alpha:
        alphas = [frame_random(w, h, seed=1300 + f, full_range=True)[0] for f in range(F)]
        abuf = torch.zeros((F, h, pitch), dtype=torch.uint8, device="cuda")
        for f, a in enumerate(alphas):
            abuf[f, :, :w] = dev(a)
    obuf = torch.full((F, h, pitch, 4), 0x5A, dtype=torch.uint8, device="cuda")
    n0 = ctx().launch_count
    ctx().decode_nv12(gamma, ybuf[:, :, :w], cbuf[:, :, :w], obuf[:, :, :w], alpha_y=abuf[:, :, :w] if alpha else None, wait=True)
    assert ctx().launch_count == n0 + 1
    got = obuf.cpu().numpy()
    assert (got[:, :, w:] == 0x5A).all(), "row padding overwritten"
    for f in range(F):
        want = ORA.decode_nv12(gamma, frames[f][0], frames[f][1], alpha_y=alphas[f] if alpha else None)
        assert np.array_equal(got[f, :, :w], want), f"{w}x{h} frame {f}: {(got[f, :, :w] != want).sum()} byte mismatches"


@pytest.mark.parametrize("w,h,F,gamma,alpha,gap", [(1920, 1080, 3, 3, False, 0), (1360, 720, 7, 2, False, 2), (720, 480, 12, 0, False, 1),
                                                  (2000, 1536, 2, 1, True, 0), (600, 64, 90, 3, False, 3), (96, 2, 9000, 1, False, 1)])
def test_strided_batches_on_the_linear_tile_stream_kernel(w, h, F, gamma, alpha, gap):
    """One allocation per frame with the CbCr plane right behind the Y plane (what a hardware decoder hands over), `gap`
    spare rows between frames, and an output whose frames are not back to back either: the batch is not one tall image.
    Ragged widths take k_decode_lin's STRIDED instantiation, which steps (frame, row pair) with a wrap (full-width rows
    stay on the round-1 ring kernel, first case; BT709_STRIDED_LIN=all moves them too, see the subprocess test).  Large
    enough for the batch kernel (more than 4 steps per warp); the last case moves whole frames per step (dq > 0)."""
    from _gpu import ctx, dev
    frames = [frame_random(w, h, seed=1500 + f, full_range=(f % 2 == 1)) for f in range(min(F, 8))]
    rows = h + h // 2 + gap
    buf = torch.zeros((F, rows, w), dtype=torch.uint8, device="cuda")
    for f in range(F):
        y, c = frames[f % len(frames)]
        buf[f, :h] = dev(y)
        buf[f, h:h + h // 2] = dev(c)
    alphas, abuf = None, None
    if alpha:
        alphas = [frame_random(w, h, seed=1600 + f, full_range=True)[0] for f in range(len(frames))]
        abuf = torch.zeros((F, h + 2, w), dtype=torch.uint8, device="cuda")
        for f in range(F):
            abuf[f, :h] = dev(alphas[f % len(frames)])
    obuf = torch.full((F, h + 1 + gap, w, 4), 0x5A, dtype=torch.uint8, device="cuda")
    n0 = ctx().launch_count
    ctx().decode_nv12(gamma, buf[:, :h], buf[:, h:h + h // 2], obuf[:, :h], alpha_y=abuf[:, :h] if alpha else None, wait=True)
    assert ctx().launch_count == n0 + 1
    assert bool((obuf[:, h:] == 0x5A).all()), "rows between the output frames overwritten"
    wants = [torch.from_numpy(ORA.decode_nv12(gamma, *frames[i], alpha_y=alphas[i] if alpha else None)).cuda() for i in range(len(frames))]
    for f in range(F):
        assert torch.equal(obuf[f, :h], wants[f % len(frames)]), f"{w}x{h} frame {f} of {F}"


def test_strided_full_width_batches_on_both_kernels_in_a_subprocess():
    """BT709_STRIDED_LIN=all: full-width strided batches on k_decode_lin's STRIDED instantiation instead of the ring kernel"""
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "tests")
from _oracle import Oracle, frame_random
from _gpu import ctx, dev
ora = Oracle()
for (w, h, F, g) in [(1920, 1080, 3, 3), (1280, 720, 7, 2), (128, 2, 9000, 1)]:
    fr = [frame_random(w, h, seed=1700 + f) for f in range(min(F, 4))]
    buf = torch.zeros((F, h + h // 2 + 1, w), dtype=torch.uint8, device="cuda")
    for f in range(F):
        buf[f, :h] = dev(fr[f % len(fr)][0]); buf[f, h:h + h // 2] = dev(fr[f % len(fr)][1])
    out = torch.zeros((F, h + 2, w, 4), dtype=torch.uint8, device="cuda")
    ctx().decode_nv12(g, buf[:, :h], buf[:, h:h + h // 2], out[:, :h], wait=True)
    wants = [torch.from_numpy(ora.decode_nv12(g, *x)).cuda() for x in fr]
    for f in range(F):
        assert torch.equal(out[f, :h], wants[f % len(fr)]), (w, h, g, f)
print("variant ok")
'''
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, BT709_STRIDED_LIN="all"), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("w,h,F,gin,gout", [(1920, 1080, 3, 1, 0), (1280, 720, 5, 1, 1), (720, 480, 7, 2, 2), (3840, 64, 6, 1, 0),
                                           (1920, 2, 40, 0, 0), (1440, 96, 9, 1, 0)])
def test_encode_linear_tile_order(w, h, F, gin, gout):
    """Rows of 15, 10, 6, 30 or 12 tiles would leave warps of a 32-warp CTA idle in the band layout; contiguous batches
    take the tiles in linear order instead (k_encode_fast, EncodeArgs.lin_warps).  Every frame against the oracle."""
    from _gpu import ctx, dev
    srcs = [bgra_random(w, h, seed=1800 + f) for f in range(min(F, 4))]
    tin = torch.stack([dev(srcs[f % len(srcs)]) for f in range(F)])
    ey = torch.zeros((F, h, w), dtype=torch.uint8, device="cuda")
    ec = torch.zeros((F, h // 2, w), dtype=torch.uint8, device="cuda")
    n0 = ctx().launch_count
    ctx().encode_bgra_to_nv12(gin, gout, tin, ey, ec, wait=True)
    assert ctx().launch_count == n0 + 1
    wants = [ORA.encode_bgra_nv12(gin, gout, s_) for s_ in srcs]
    gy, gc = ey.cpu().numpy(), ec.cpu().numpy()
    for f in range(F):
        wy, wc = wants[f % len(srcs)]
        assert np.array_equal(gy[f], wy) and np.array_equal(gc[f], wc), f"{w}x{h} frame {f}"
