"""The reference's own XCTest cases, replayed through the host layer that mirrors its classes
(include/metal_bt709_decoder.h).  Structure follows MetalBT709DecoderTests.m:189-277
(-convert_bt709_to_srgb:type:) and AppleEncodeDecodeBT709Tests.m.  -m gpu."""
import json
import os

import numpy as np
import pytest

from _oracle import Oracle, bgra_random, frame_random

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
ORA = Oracle()


def R():
    from metalbt709decoder_b200 import renderer
    return renderer


@pytest.fixture(scope="module")
def mrc():
    m = R().MetalRenderContext(0)
    assert m.setupMetal()
    yield m
    m.release()


def rgbToPixel(r, g, b):
    return (0xFF << 24) | (r << 16) | (g << 8) | b


def pixelToRGB(p):
    return (int(p) >> 16) & 0xFF, (int(p) >> 8) & 0xFF, int(p) & 0xFF


def convert_bt709_to_srgb(mrc, bt709, type):
    """MetalBT709DecoderTests.m:189-277"""
    r = R()
    width, height = 2, 2
    outBT709 = np.full(width * height, bt709, np.uint32)
    if type in (r.BGRAToBT709ConverterSoftware, r.BGRAToBT709ConverterVImage):
        worked, outBGRA = r.BGRAToBT709Converter.unconvert(mrc, outBT709, width, height, type)
        assert worked
        return outBGRA[0]
    metalDecoder = r.MetalBT709Decoder()
    metalDecoder.metalRenderContext = mrc
    assert metalDecoder.setupMetal()
    bgraSRGBTexture = mrc.makeBGRATexture(width, height, None, isSRGB=True)
    yCbCrBuffer = r.BGRAToBT709Converter.createCoreVideoYCbCrBuffer(width, height)
    r.BGRAToBT709Converter.setBT709Attributes(yCbCrBuffer)
    assert r.BGRAToBT709Converter.copyBT709ToCoreVideo(mrc, outBT709, yCbCrBuffer)
    worked = metalDecoder.decodeBT709(yCbCrBuffer, None, bgraSRGBTexture, None, width, height, True)
    assert worked, "decodeBT709"
    return mrc.getBGRATexturePixels(bgraSRGBTexture)[0, 0]


METAL = VEC["files"]["MetalBT709DecoderTests.m"]


@pytest.mark.parametrize("t", METAL, ids=[t["name"] for t in METAL])
def test_MetalBT709DecoderTests(mrc, t):
    c = {k["var"]: k["expected"] for k in t["checks"]}
    yuv = (c["Cr"] << 16) | (c["Cb"] << 8) | c["Y"]  # the vImage encoder's asserted output
    out = convert_bt709_to_srgb(mrc, yuv, R().BGRAToBT709ConverterMetal)
    assert pixelToRGB(out) == (c["Rout"], c["Gout"], c["Bout"])
    assert (int(out) >> 24) == 0xFF


APPLE_SW = [t for t in VEC["files"]["AppleEncodeDecodeBT709Tests.m"]
            if t["backends"].get("encodeType") == "BGRAToBT709ConverterSoftware"]


@pytest.mark.parametrize("t", APPLE_SW, ids=[t["name"] for t in APPLE_SW])
def test_AppleEncodeDecodeBT709Tests_software(mrc, t):
    r = R()
    c, i = {k["var"]: k["expected"] for k in t["checks"]}, t["inputs"]
    src = np.full(4, rgbToPixel(i["Rin"], i["Gin"], i["Bin"]), np.uint32)
    worked, yuv = r.BGRAToBT709Converter.convert(mrc, src, 2, 2, r.BGRAToBT709ConverterSoftware)
    assert worked
    assert (int(yuv[0]) & 0xFF, (int(yuv[0]) >> 8) & 0xFF, (int(yuv[0]) >> 16) & 0xFF) == (c["Y"], c["Cb"], c["Cr"])
    dec = {"BGRAToBT709ConverterSoftware": r.BGRAToBT709ConverterSoftware,
           "BGRAToBT709ConverterMetal": r.BGRAToBT709ConverterMetal}[t["backends"]["decodeType"]]
    out = convert_bt709_to_srgb(mrc, yuv[0], dec)
    assert pixelToRGB(out) == (c["Rout"], c["Gout"], c["Bout"])


def _buffer(w, h, y, c, tf):
    r = R()
    pb = r.BGRAToBT709Converter.createCoreVideoYCbCrBuffer(w, h)
    pb.y[:] = y
    pb.cbcr[:] = c
    pb.setAttachments(r.kMBDYCbCrMatrix_ITU_R_709_2, tf)
    return pb


def test_decoder_guard_clauses(mrc):
    """-processBT709ToSRGB: returns FALSE on these (MetalBT709Decoder.m:272-368)"""
    r = R()
    w, h = 16, 8
    y, c = frame_random(w, h)
    d = r.MetalBT709Decoder()
    d.metalRenderContext = mrc
    tex = mrc.makeBGRATexture(w, h)
    pb = _buffer(w, h, y, c, r.kMBDTransferFunction_ITU_R_709_2)
    assert d.decodeBT709(pb, None, tex, None, w, h, True)
    assert not d.decodeBT709(pb, None, tex, None, w + 2, h, True)            # render size mismatch
    assert not d.decodeBT709(pb, None, mrc.makeBGRATexture(w, h + 2), None, w, h, True)  # texture size mismatch
    pb.setAttachments(r.kMBDYCbCrMatrix_ITU_R_601_4, r.kMBDTransferFunction_ITU_R_709_2)
    assert not d.decodeBT709(pb, None, tex, None, w, h, True)                # only the BT.709 matrix is supported
    pb.setAttachments(r.kMBDYCbCrMatrix_ITU_R_709_2, r.kMBDTransferFunction_sRGB)
    assert not d.decodeBT709(pb, None, tex, None, w, h, True)                # gamma Apple needs the 709 tag
    d2 = r.MetalBT709Decoder()
    d2.metalRenderContext = mrc
    d2.gamma = r.MetalBT709GammaSRGB
    assert d2.decodeBT709(pb, None, tex, None, w, h, True)
    assert np.array_equal(tex.pixels, ORA.decode_nv12(1, y, c))
    d3 = r.MetalBT709Decoder()
    d3.metalRenderContext = mrc
    d3.setUseComputeRenderer(True)
    d3.setHasAlphaChannel(True)
    assert not d3.setupMetal()  # MetalBT709Decoder.m:108


def test_alpha_decoder(mrc):
    """hasAlphaChannel forces sRGB gamma (MetalBT709Decoder.m:165-169); alpha stream must be tagged linear"""
    r = R()
    w, h = 64, 16
    y, c = frame_random(w, h, seed=1)
    ay, ac = frame_random(w, h, seed=2)
    d = r.MetalBT709Decoder()
    d.metalRenderContext = mrc
    d.setHasAlphaChannel(True)
    assert d.setupMetal() and d.gamma == r.MetalBT709GammaSRGB
    tex = mrc.makeBGRATexture(w, h)
    rgb = _buffer(w, h, y, c, r.kMBDTransferFunction_sRGB)
    alpha = _buffer(w, h, ay, ac, r.kMBDTransferFunction_sRGB)
    assert not d.decodeBT709(rgb, alpha, tex, None, w, h, True)  # :357-368
    alpha.setAttachments(r.kMBDYCbCrMatrix_ITU_R_709_2, r.kMBDTransferFunction_Linear)
    assert d.decodeBT709(rgb, alpha, tex, None, w, h, True)
    assert np.array_equal(tex.pixels, ORA.decode_nv12(1, y, c, alpha_y=ay))


def test_async_decodes_overlap_and_complete(mrc):
    """waitUntilCompleted:FALSE queues work; -waitUntilCompleted drains the ring"""
    r = R()
    w, h, n = 320, 180, 9
    d = r.MetalBT709Decoder()
    d.metalRenderContext = mrc
    d.gamma = r.MetalBT709GammaLinear
    frames = [frame_random(w, h, seed=100 + i) for i in range(n)]
    bufs = [_buffer(w, h, y, c, r.kMBDTransferFunction_Linear) for y, c in frames]
    texs = [mrc.makeBGRATexture(w, h) for _ in range(n)]
    n0 = mrc.launchCount
    for pb, t in zip(bufs, texs):
        assert d.decodeBT709(pb, None, t, None, w, h, False)
    assert mrc.waitUntilCompleted()
    assert mrc.launchCount == n0 + n
    for (y, c), t in zip(frames, texs):
        assert np.array_equal(t.pixels, ORA.decode_nv12(2, y, c))


def test_renderScaled_and_AAPLRenderer_policy(mrc):
    r = R()
    w, h, ow, oh = 128, 72, 171, 96
    y, c = frame_random(w, h, seed=5)
    pb = _buffer(w, h, y, c, r.kMBDTransferFunction_sRGB)
    mid_want = ORA.decode_nv12(1, y, c)
    ren = r.AAPLRenderer(mrc, r.MetalBT709GammaSRGB)
    same = mrc.makeBGRATexture(w, h)
    n0 = mrc.launchCount
    assert ren.drawFrame(pb, None, same)                     # exact size: one pass (AAPLRenderer.m:905-939)
    assert mrc.launchCount == n0 + 1
    assert np.array_equal(same.pixels, mid_want)
    big = mrc.makeBGRATexture(ow, oh)
    assert ren.drawFrame(pb, None, big)                      # two passes (AAPLRenderer.m:940-977)
    assert mrc.launchCount == n0 + 3
    want = ORA.rescale_bgra(mid_want, ow, oh)
    assert np.array_equal(big.pixels, want)
    sc = r.MetalScaleRenderContext()
    big2 = mrc.makeBGRATexture(ow, oh)
    assert sc.renderScaled(mrc, big2, ow, oh, None, same)
    assert np.array_equal(big2.pixels, want)
    assert not sc.renderScaled(mrc, big2, 0, oh, None, same)


def test_createYCbCrFromBGRA_gamma_policy(mrc):
    """+createYCbCrFromCGImage:isLinear:asSRGBGamma: picks the gamma pair (BGRAToBT709Converter.m:919-937)"""
    r = R()
    w, h = 64, 32
    bgra = bgra_random(w, h, seed=711)
    words = bgra.view(np.uint32).reshape(h, w)
    for kw, pair, tf in [({}, (1, 0), r.kMBDTransferFunction_ITU_R_709_2),
                         ({"asSRGBGamma": True}, (1, 1), r.kMBDTransferFunction_sRGB),
                         ({"isLinear": True}, (2, 2), r.kMBDTransferFunction_Linear)]:
        pb = r.BGRAToBT709Converter.createYCbCrFromBGRA(mrc, words, w, h, **kw)
        assert pb is not None and pb.struct.transferFunction == tf
        wy, wc = ORA.encode_bgra_nv12(pair[0], pair[1], bgra)
        assert np.array_equal(pb.y, wy) and np.array_equal(pb.cbcr, wc)
    assert r.BGRAToBT709Converter.createYCbCrFromBGRA(mrc, words, w, h, isLinear=True, asSRGBGamma=True) is None
    dst = r.BGRAToBT709Converter.createCoreVideoYCbCrBuffer(w, h)
    assert r.cvpbu_ycbcr_subsample(mrc, words, w, h, dst, r.BT709GammaApple, r.BT709GammaSrgb)
    wy, wc = ORA.encode_bgra_nv12(0, 1, bgra)
    assert np.array_equal(dst.y, wy) and np.array_equal(dst.cbcr, wc)
    assert not r.cvpbu_ycbcr_subsample(mrc, words, w, h, dst, 9, r.BT709GammaSrgb)
