"""BASELINE.json configurations at their full sizes (-m gpu).  The oracle computes a handful of distinct frames;
the batch cycles through them, so every one of the hundreds of output frames has a known answer (size-independent
property: frames are independent, a batch is the per-frame result repeated)."""
import numpy as np
import pytest
import torch

from _oracle import GAMMA_APPLE, GAMMA_SRGB, Oracle, bgra_random, frame_random

pytestmark = pytest.mark.gpu
ORA = Oracle()
K = 3  # distinct frames


def _cycled(tensors, frames):
    t = torch.stack(tensors)
    return t.repeat((frames + K - 1) // K, *([1] * (t.dim() - 1)))[:frames].contiguous()


def _assert_cycled(out, wants, what):
    frames = out.shape[0]
    for k in range(K):
        want = torch.from_numpy(wants[k]).cuda()
        sel = out[k::K]
        bad = (sel != want[None]).flatten(1).any(dim=1)
        assert not bool(bad.any()), f"{what}: frames {[k + K * int(i) for i in bad.nonzero()[:5].flatten()]} differ"
    assert frames == out.shape[0]


def test_c2_600_frames_1080p_decode_then_rescale():
    """configs[1]: 600 x 1920x1080 decoded (BT.709 curve) in one call, then pass 2 to 2560x1440 on 200 of them"""
    from _gpu import ctx, dev
    w, h, F = 1920, 1080, 600
    src = [frame_random(w, h, seed=900 + k) for k in range(K)]
    ty = _cycled([dev(y) for y, _ in src], F)
    tc = _cycled([dev(c) for _, c in src], F)
    out = torch.zeros((F, h, w, 4), dtype=torch.uint8, device="cuda")
    ctx().decode_nv12(3, ty, tc, out, wait=True)
    wants = [ORA.decode_nv12(3, y, c) for y, c in src]
    _assert_cycled(out, wants, "C2 decode")
    F2 = 200
    big = torch.zeros((F2, 1440, 2560, 4), dtype=torch.uint8, device="cuda")
    ctx().rescale_bgra(out[:F2], big, wait=True)
    _assert_cycled(big, [ORA.rescale_bgra(wd, 2560, 1440) for wd in wants], "C2 pass 2")


def test_c3_uhd_batch_decode():
    """configs[2] per-GPU shape: UHD frames in one call (96 here; the bench runs 300 per GPU)"""
    from _gpu import ctx, dev
    w, h, F = 3840, 2160, 96
    src = [frame_random(w, h, seed=910 + k) for k in range(K)]
    ty = _cycled([dev(y) for y, _ in src], F)
    tc = _cycled([dev(c) for _, c in src], F)
    out = torch.zeros((F, h, w, 4), dtype=torch.uint8, device="cuda")
    ctx().decode_nv12(0, ty, tc, out, wait=True)
    _assert_cycled(out, [ORA.decode_nv12(0, y, c) for y, c in src], "C3 decode")


def test_c4_alpha_and_c5_encode_batches():
    """configs[3] and [4]: 2048x1536 RGB + alpha streams decoded (sRGB gamma), and the sRGB -> Apple encode"""
    from _gpu import ctx, dev
    w, h, F = 2048, 1536, 120
    src = [frame_random(w, h, seed=920 + k) for k in range(K)]
    alp = [frame_random(w, h, seed=930 + k, full_range=True)[0] for k in range(K)]
    ty = _cycled([dev(y) for y, _ in src], F)
    tc = _cycled([dev(c) for _, c in src], F)
    ta = _cycled([dev(a) for a in alp], F)
    out = torch.zeros((F, h, w, 4), dtype=torch.uint8, device="cuda")
    ctx().decode_nv12(GAMMA_SRGB, ty, tc, out, alpha_y=ta, wait=True)
    _assert_cycled(out, [ORA.decode_nv12(GAMMA_SRGB, src[k][0], src[k][1], alpha_y=alp[k]) for k in range(K)], "C4")
    imgs = [bgra_random(w, h, seed=940 + k) for k in range(K)]
    tin = _cycled([dev(i) for i in imgs], F)
    ey = torch.zeros((F, h, w), dtype=torch.uint8, device="cuda")
    ec = torch.zeros((F, h // 2, w), dtype=torch.uint8, device="cuda")
    ctx().encode_bgra_to_nv12(GAMMA_SRGB, GAMMA_APPLE, tin, ey, ec, wait=True)
    enc = [ORA.encode_bgra_nv12(GAMMA_SRGB, GAMMA_APPLE, i) for i in imgs]
    _assert_cycled(ey, [e[0] for e in enc], "C5 Y")
    _assert_cycled(ec, [e[1] for e in enc], "C5 CbCr")
