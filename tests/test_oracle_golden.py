"""Pin the CPU oracle against the reference's own known-answer tests.

Vectors: tests/golden/reference_vectors.json, harvested from
/root/reference/EmptyiOSTests/*.m by tests/golden/harvest_golden.py (committed).
"""
import json
import os

import pytest

from _oracle import GAMMA_APPLE, GAMMA_BT709, GAMMA_LINEAR, GAMMA_SRGB, Oracle, ref_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))

REF_GAMMA = {"BT709GammaSrgb": GAMMA_SRGB, "BT709GammaApple": GAMMA_APPLE, "BT709GammaLinear": GAMMA_LINEAR}


def _impls():
    out = [pytest.param(Oracle(), id="restatement")]
    r = ref_oracle()
    if r is not None:
        out.append(pytest.param(r, id="reference_ref"))
    return out


IMPLS = _impls()


def _chk(t):
    return {c["var"]: c["expected"] for c in t["checks"]}


# ---- MetalBT709DecoderTests.m: (Y,Cb,Cr) -> Metal decode with the default (Apple) gamma -------
METAL = VEC["files"]["MetalBT709DecoderTests.m"]


def test_harvest_counts():
    # 28 live Metal-decode tests, 23 Apple encode/decode tests, 22 pure-C tests
    assert len(METAL) == 28
    assert len(VEC["files"]["AppleEncodeDecodeBT709Tests.m"]) == 23
    assert len(VEC["files"]["CoreImageMetalFilterTests.m"]) == 22


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("t", METAL, ids=[t["name"] for t in METAL])
def test_metal_decode_goldens(impl, t):
    c = _chk(t)
    # the encode half of these tests is vImage (closed source); its asserted output is our input
    got = impl.decode_pixel(GAMMA_APPLE, c["Y"], c["Cb"], c["Cr"])
    assert got == (c["Rout"], c["Gout"], c["Bout"]), (t["name"], t["line"])


# ---- AppleEncodeDecodeBT709Tests.m -------------------------------------------------------------
APPLE = VEC["files"]["AppleEncodeDecodeBT709Tests.m"]
APPLE_SW = [t for t in APPLE if t["backends"].get("encodeType") == "BGRAToBT709ConverterSoftware"]
APPLE_DEC = [t for t in APPLE if t["backends"].get("decodeType") in
             ("BGRAToBT709ConverterSoftware", "BGRAToBT709ConverterMetal")]


def test_apple_software_count():
    assert len(APPLE_SW) == 14


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("t", APPLE_SW, ids=[t["name"] for t in APPLE_SW])
def test_apple_software_encode_goldens(impl, t):
    c, i = _chk(t), t["inputs"]
    assert impl.encode444_pixel(GAMMA_APPLE, i["Rin"], i["Gin"], i["Bin"]) == (c["Y"], c["Cb"], c["Cr"])


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("t", APPLE_DEC, ids=[t["name"] for t in APPLE_DEC])
def test_apple_decode_goldens(impl, t):
    c = _chk(t)
    assert impl.decode_pixel(GAMMA_APPLE, c["Y"], c["Cb"], c["Cr"]) == (c["Rout"], c["Gout"], c["Bout"])


# ---- CoreImageMetalFilterTests.m: pure BT709.h functions ----------------------------------------
CI = {t["name"]: t for t in VEC["files"]["CoreImageMetalFilterTests.m"]}


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", [n for n in CI if n.startswith("testConvertsLinearRGBToBT709_") and "srgb" not in n])
def test_linear_rgb_to_bt709(impl, name):
    t = CI[name]
    c, i = _chk(t), t["inputs"]
    ag = i["applyGammaMap"]
    assert impl.linear_rgb_to_ycbcr(i["Rin"], i["Gin"], i["Bin"], ag) == (c["Y"], c["Cb"], c["Cr"])
    assert impl.ycbcr_to_linear_rgb(c["Y"], c["Cb"], c["Cr"], ag) == (c["R"], c["G"], c["B"])


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", ["testConvertsLinearRGBToBT709_42PercentBlue_srgb",
                                  "testConvertsSRGBToYCbCr_ResampleUpperNW", "testConvertsSRGBToYCbCr_ResampleUpperNE",
                                  "testConvertsSRGBToYCbCr_ResampleUpperSW", "testConvertsSRGBToYCbCr_ResampleUpperSE"])
def test_bt709_spec_curve_pixels(impl, name):
    t = CI[name]
    c, i = _chk(t), t["inputs"]
    assert "BT709_from_sRGB_convertRGBToYCbCr" in t["calls"] and "BT709_to_sRGB_convertYCbCrToRGB" in t["calls"]
    assert impl.encode444_pixel(GAMMA_BT709, i["Rin"], i["Gin"], i["Bin"]) == (c["Y"], c["Cb"], c["Cr"])
    assert impl.decode_pixel(GAMMA_BT709, c["Y"], c["Cb"], c["Cr"]) == (c["R"], c["G"], c["B"])


AVG = [n for n in CI if n.startswith("testConvertsSRGBToYCbCr_AverageOf4")]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", AVG)
def test_average_of_4(impl, name):
    t = CI[name]
    c, i = _chk(t), t["inputs"]
    rgb = [i[f"{ch}{k}"] for k in (1, 2, 3, 4) for ch in "RGB"]
    gin, gout = REF_GAMMA[t["gammas"]["inputGamma"]], REF_GAMMA[t["gammas"]["outputGamma"]]
    y4, cb, cr = impl.average_block(rgb, gin, gout)
    assert y4 == [c["Y1"], c["Y2"], c["Y3"], c["Y4"]]
    assert (cb, cr) == (c["Cb"], c["Cr"])


@pytest.mark.parametrize("impl", IMPLS)
def test_srgb_average_cbcr(impl):
    t = CI["testConvertsSRGBToYCbCr_Average"]
    c, i = _chk(t), t["inputs"]
    rgb = [i[f"{ch}{k}"] for k in (1, 2, 3, 4) for ch in "RGB"]
    out = impl.srgb_average_cbcr(rgb)
    for k in range(4):
        got = impl.encode444_pixel(GAMMA_SRGB, *out[k * 3:k * 3 + 3])
        assert got == (c[f"Y{k+1}"], c[f"Cb{k+1}"], c[f"Cr{k+1}"])


@pytest.mark.parametrize("impl", IMPLS[:1])
@pytest.mark.parametrize("key", ["bt709", "apple", "srgb"])
def test_roundtrip_histograms(impl, key):
    h = VEC["histograms"][key]
    got = impl.roundtrip_histogram(h["gamma"])
    assert got[0] == h["exact"]
    assert got[:len(h["buckets"])] == h["buckets"]
    assert sum(got) == 1 << 24 and sum(got[len(h["buckets"]):]) == 0


def test_alpha_plane_against_the_metal_shader_arithmetic():
    """D12 has no test and no exact CPU twin in the reference ("parity unpinned").  What can be pinned on the CPU: the
    shader's own float32 arithmetic (AAPLShaders.metal:249-271: (A/255 - 16/255) * 1.1644, saturate, unorm8 store) gives
    the same byte as the oracle (the CPU near-twin BT709.h:466-513, which multiplies by 255/219) for ALL 256 codes, with
    round-half-even or half-away alike: the literal 1.1644 never matters (the pre-rounding value stays 0.0044 away from
    a .5).  If the R8Unorm sample is really delivered as a `half` (the shader declares texture2d<half>), six codes come
    out one LSB off - the bound of what real Metal hardware can differ from this implementation by."""
    import numpy as np
    f32 = np.float32
    ora = Oracle()
    ours = np.array([ora.alpha_pixel(a) for a in range(256)])

    def shader(sample):
        yn = (sample - f32(16.0) / f32(255.0)).astype(np.float32)
        yn = np.clip((yn * f32(1.1644)).astype(np.float32), 0, 1).astype(np.float32)
        return (yn * f32(255.0)).astype(np.float32)

    s32 = (np.arange(256).astype(np.float32) / f32(255.0)).astype(np.float32)
    x = shader(s32)
    assert np.array_equal(np.rint(x).astype(int), ours) and np.array_equal(np.floor(x + 0.5).astype(int), ours)
    assert np.abs((x - np.floor(x)) - 0.5).min() > 0.004
    xh = shader(s32.astype(np.float16).astype(np.float32))
    d = np.rint(xh).astype(int) - ours
    assert np.abs(d).max() == 1 and np.nonzero(d)[0].tolist() == [86, 92, 104, 141, 177, 232]
