"""srgb_to_bt709 command-line tool end to end (PNG -> GPU 4:2:0 encode -> .y4m), and the way back
(.y4m planes -> NV12 on the GPU -> decode).  -m gpu."""
import os
import subprocess

import numpy as np
import pytest
import torch

from _oracle import GAMMA_APPLE, GAMMA_LINEAR, GAMMA_SRGB, Oracle
from metalbt709decoder_b200 import _lib

pytestmark = pytest.mark.gpu
ORA = Oracle()
CLI = os.path.join(os.path.dirname(_lib.LIB_PATH), "srgb_to_bt709")


def _png(path, rgba):
    from PIL import Image
    Image.fromarray(rgba, "RGBA" if rgba.shape[2] == 4 else "RGB").save(path)


def _premultiplied_bgra(rgba):
    a = rgba[..., 3].astype(np.uint32) if rgba.shape[2] == 4 else np.full(rgba.shape[:2], 255, np.uint32)
    out = np.empty(rgba.shape[:2] + (4,), np.uint8)
    for dst, src in ((0, 2), (1, 1), (2, 0)):
        out[..., dst] = (rgba[..., src].astype(np.uint32) * a + 127) // 255
    out[..., 3] = 255
    return out, a.astype(np.uint8)


def _read_y4m(path, expect_frames):
    raw = open(path, "rb").read()
    head, rest = raw.split(b"\n", 1)
    assert rest.startswith(b"XYSCSS=420JPEG\n")
    rest = rest[15:]
    tok = dict((t[:1], t[1:]) for t in head.split()[1:])
    w, h = int(tok[b"W"]), int(tok[b"H"])
    frames = []
    fsz = 6 + w * h * 3 // 2
    assert len(rest) == fsz * expect_frames
    for i in range(expect_frames):
        f = rest[i * fsz:(i + 1) * fsz]
        assert f[:6] == b"FRAME\n"
        y = np.frombuffer(f, np.uint8, w * h, 6).reshape(h, w)
        u = np.frombuffer(f, np.uint8, w * h // 4, 6 + w * h).reshape(h // 2, w // 2)
        v = np.frombuffer(f, np.uint8, w * h // 4, 6 + w * h * 5 // 4).reshape(h // 2, w // 2)
        frames.append((y, u, v))
    return head, frames


@pytest.mark.parametrize("gamma,pair", [("apple", (GAMMA_SRGB, GAMMA_APPLE)), ("srgb", (GAMMA_SRGB, GAMMA_SRGB)),
                                        ("linear", (GAMMA_LINEAR, GAMMA_LINEAR))])
def test_cli_frame_with_alpha(tmp_path, gamma, pair):
    rng = np.random.default_rng(3)
    rgba = rng.integers(0, 256, (48, 64, 4), dtype=np.uint8)
    png, out = str(tmp_path / "in.png"), str(tmp_path / "out.y4m")
    _png(png, rgba)
    r = subprocess.run([CLI, "-gamma", gamma, "-alpha", "1", "-frame", png, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"wrote {out}" in r.stdout and "out_alpha.y4m" in r.stdout
    bgra, a = _premultiplied_bgra(rgba)
    head, frames = _read_y4m(out, 1)
    assert head == b"YUV4MPEG2 W64 H48 F1:1 Ip A1:1 C420jpeg"  # -frame defaults to 1 fps (srgb_to_bt709.m:1298-1299)
    wy, wc = ORA.encode_bgra_nv12(pair[0], pair[1], bgra)
    y, u, v = frames[0]
    assert np.array_equal(y, wy) and np.array_equal(u, wc[:, 0::2]) and np.array_equal(v, wc[:, 1::2])
    # the alpha stream: A as gray, always linear (srgb_to_bt709.m:842-954)
    gray = np.stack([a, a, a, np.full_like(a, 255)], axis=-1)
    _, aframes = _read_y4m(str(tmp_path / "out_alpha.y4m"), 1)
    ay, ac = ORA.encode_bgra_nv12(GAMMA_LINEAR, GAMMA_LINEAR, gray)
    assert np.array_equal(aframes[0][0], ay) and np.array_equal(aframes[0][1], ac[:, 0::2])


def test_cli_frames_pattern_and_decode_back(tmp_path):
    """-frames F0001.png picks up consecutive files; the y4m planes re-interleaved on the GPU decode
    to what the oracle decodes from the oracle's own encode"""
    from _gpu import ctx, dev
    from metalbt709decoder_b200._lib import Plane
    import ctypes
    rng = np.random.default_rng(4)
    imgs = [rng.integers(0, 256, (32, 128, 3), dtype=np.uint8) for _ in range(3)]
    for i, im in enumerate(imgs):
        _png(str(tmp_path / f"F{i + 7:04d}.png"), im)
    out = str(tmp_path / "clip.y4m")
    r = subprocess.run([CLI, "-gamma", "srgb", "-frames", str(tmp_path / "F0007.png"), out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    head, frames = _read_y4m(out, 3)
    assert b"F30:1" in head  # -frames defaults to 30 fps
    C, L = ctx(), _lib.lib()
    for im, (y, u, v) in zip(imgs, frames):
        bgra, _ = _premultiplied_bgra(im)
        wy, wc = ORA.encode_bgra_nv12(GAMMA_SRGB, GAMMA_SRGB, bgra)
        assert np.array_equal(y, wy)
        tu, tv = dev(u)[None], dev(v)[None]
        tc = torch.zeros((1, 16, 128), dtype=torch.uint8, device="cuda")
        pu, pv, pc = Plane(tu.data_ptr(), 64, 0), Plane(tv.data_ptr(), 64, 0), Plane(tc.data_ptr(), 128, 0)
        assert L.bt709_planar_to_cbcr(C.handle, ctypes.byref(pu), ctypes.byref(pv), ctypes.byref(pc), 64, 16, 1, None, 1) == 0
        assert np.array_equal(tc[0].cpu().numpy(), wc)
        outt = torch.zeros((1, 32, 128, 4), dtype=torch.uint8, device="cuda")
        C.decode_nv12(1, dev(y)[None], tc, outt, wait=True)
        assert np.array_equal(outt[0].cpu().numpy(), ORA.decode_nv12(1, wy, wc))
        # and the split is the inverse of the merge
        tu2, tv2 = torch.zeros_like(tu), torch.zeros_like(tv)
        pu2, pv2 = Plane(tu2.data_ptr(), 64, 0), Plane(tv2.data_ptr(), 64, 0)
        assert L.bt709_cbcr_to_planar(C.handle, ctypes.byref(pc), ctypes.byref(pu2), ctypes.byref(pv2), 64, 16, 1, None, 1) == 0
        assert torch.equal(tu2, tu) and torch.equal(tv2, tv)


def test_cli_rejects_odd_sizes(tmp_path):
    png = str(tmp_path / "odd.png")
    _png(png, np.zeros((5, 8, 3), np.uint8))
    r = subprocess.run([CLI, "-frame", png, str(tmp_path / "o.y4m")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "must be even" in r.stderr
