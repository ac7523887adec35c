"""Validate the restatement (oracle/oracle_bt709.c) against the reference's own
headers compiled where they lie (oracle/_ref).  Skipped when _ref is unavailable."""
import numpy as np
import pytest

from _oracle import (GAMMA_APPLE, GAMMA_LINEAR, GAMMA_SRGB, Oracle, bgra_random, frame_lattice,
                     frame_random, ref_oracle)

REF = ref_oracle()
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (needs /root/reference)")
ORA = Oracle()


@pytest.mark.parametrize("gamma", [0, 1, 2, 3])
def test_decode_all_2_24_inputs(gamma):
    y, c = frame_lattice()
    a = ORA.decode_nv12(gamma, y, c)
    b = REF.decode_nv12(gamma, y, c)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("gin,gout", [(GAMMA_SRGB, GAMMA_APPLE), (GAMMA_SRGB, GAMMA_SRGB), (GAMMA_LINEAR, GAMMA_LINEAR),
                                      (GAMMA_APPLE, GAMMA_SRGB), (GAMMA_SRGB, GAMMA_LINEAR), (GAMMA_APPLE, GAMMA_APPLE)])
def test_encode420_random(gin, gout):
    bgra = bgra_random(256, 128, seed=711)
    ya, ca = ORA.encode_bgra_nv12(gin, gout, bgra)
    yb, cb = REF.encode_bgra_nv12(gin, gout, bgra)
    assert np.array_equal(ya, yb) and np.array_equal(ca, cb)


@pytest.mark.parametrize("gamma", [0, 1, 2, 3])
def test_encode444_sampled(gamma):
    rng = np.random.default_rng(5)
    for r, g, b in rng.integers(0, 256, (4000, 3)):
        assert ORA.encode444_pixel(gamma, int(r), int(g), int(b)) == REF.encode444_pixel(gamma, int(r), int(g), int(b))


def test_alpha_all_256():
    assert [ORA.alpha_pixel(a) for a in range(256)] == [REF.alpha_pixel(a) for a in range(256)]


def test_srgb_curves_match():
    xs = np.linspace(0, 1, 20001, dtype=np.float32)
    for x in xs:
        assert ORA.lib.oracle_srgb_to_linear(float(x)) == REF.lib.ref_srgb_to_linear(float(x))
        assert ORA.lib.oracle_linear_to_srgb(float(x)) == REF.lib.ref_linear_to_srgb(float(x))


def test_odd_sized_decode():
    y, c = frame_random(37, 21, seed=3, full_range=True)
    for gamma in range(4):
        assert np.array_equal(ORA.decode_nv12(gamma, y, c), REF.decode_nv12(gamma, y, c))
