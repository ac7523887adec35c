"""YUV4MPEG2 writer/reader (include/y4m_b200.h) and the srgb_to_bt709 CLI's argument handling. CPU only.

The writer must be byte-identical to the reference's Renderer/y4m_writer.h (which needs NSString
and cannot be compiled here): header text from y4m_writer.h:61-192, frame layout from :194-241."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from metalbt709decoder_b200 import _lib

libc = ctypes.CDLL(None)
libc.fopen.restype = ctypes.c_void_p
libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
libc.fclose.argtypes = [ctypes.c_void_p]


class Hdr(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("fps", ctypes.c_int)]


class Frame(ctypes.Structure):
    _fields_ = [("yPtr", ctypes.c_void_p), ("yLen", ctypes.c_int), ("uPtr", ctypes.c_void_p), ("uLen", ctypes.c_int),
                ("vPtr", ctypes.c_void_p), ("vLen", ctypes.c_int)]


FPS = {0: b"F1:1 ", 1: b"F15:1 ", 2: b"F24:1 ", 3: b"F25:1 ", 4: b"F30000:1001 ", 5: b"F30:1 ", 6: b"F60:1 "}


def _write(path, w, h, fps, frames):
    L = _lib.lib()
    f = L.y4m_open_file(path.encode())
    assert f
    hs = Hdr(w, h, fps)
    assert L.y4m_write_header(f, ctypes.byref(hs)) == 0
    for y, u, v in frames:
        fs = Frame(y.ctypes.data, y.size, u.ctypes.data, u.size, v.ctypes.data, v.size)
        assert L.y4m_write_frame(f, ctypes.byref(fs)) == 0
    libc.fclose(f)


@pytest.mark.parametrize("fps", sorted(FPS))
def test_header_bytes_match_reference_writer(tmp_path, fps):
    p = str(tmp_path / "h.y4m")
    _write(p, 1920, 1080, fps, [])
    want = b"YUV4MPEG2 W1920 H1080 " + FPS[fps] + b"Ip A1:1 C420jpeg\nXYSCSS=420JPEG\n"
    assert open(p, "rb").read() == want


def test_bad_fps_is_rejected_like_the_reference(tmp_path):
    L = _lib.lib()
    f = L.y4m_open_file(str(tmp_path / "x.y4m").encode())
    assert L.y4m_write_header(f, ctypes.byref(Hdr(4, 4, 99))) == 3  # y4m_writer.h:141-145
    libc.fclose(f)


def test_frames_roundtrip(tmp_path):
    L = _lib.lib()
    rng = np.random.default_rng(1)
    w, h = 6, 4
    frames = [(rng.integers(0, 256, w * h, dtype=np.uint8), rng.integers(0, 256, w * h // 4, dtype=np.uint8),
               rng.integers(0, 256, w * h // 4, dtype=np.uint8)) for _ in range(3)]
    p = str(tmp_path / "f.y4m")
    _write(p, w, h, 5, frames)
    raw = open(p, "rb").read()
    head = b"YUV4MPEG2 W6 H4 F30:1 Ip A1:1 C420jpeg\nXYSCSS=420JPEG\n"
    body = b"".join(b"FRAME\n" + y.tobytes() + u.tobytes() + v.tobytes() for y, u, v in frames)
    assert raw == head + body
    f = libc.fopen(p.encode(), b"rb")
    hs = Hdr()
    assert L.y4m_read_header(f, ctypes.byref(hs)) == 0
    assert (hs.width, hs.height, hs.fps) == (w, h, 5)
    for y, u, v in frames:
        yy, uu, vv = np.zeros_like(y), np.zeros_like(u), np.zeros_like(v)
        assert L.y4m_read_frame(f, ctypes.byref(hs), yy.ctypes.data, uu.ctypes.data, vv.ctypes.data) == 0
        assert np.array_equal(yy, y) and np.array_equal(uu, u) and np.array_equal(vv, v)
    dummy = np.zeros(w * h, np.uint8)
    assert L.y4m_read_frame(f, ctypes.byref(hs), dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data) == 1  # EOF
    libc.fclose(f)


def test_reader_accepts_plain_y4m_without_the_comment_line(tmp_path):
    L = _lib.lib()
    p = str(tmp_path / "g.y4m")
    open(p, "wb").write(b"YUV4MPEG2 W2 H2 F25:1 Ip A1:1 C420mpeg2\nFRAME\n" + bytes(range(6)))
    f = libc.fopen(p.encode(), b"rb")
    hs = Hdr()
    assert L.y4m_read_header(f, ctypes.byref(hs)) == 0 and (hs.width, hs.height, hs.fps) == (2, 2, 3)
    y, u, v = np.zeros(4, np.uint8), np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    assert L.y4m_read_frame(f, ctypes.byref(hs), y.ctypes.data, u.ctypes.data, v.ctypes.data) == 0
    assert y.tolist() == [0, 1, 2, 3] and u[0] == 4 and v[0] == 5
    libc.fclose(f)


CLI = os.path.join(os.path.dirname(_lib.LIB_PATH), "srgb_to_bt709")


def _cli(*args):
    _lib.lib()  # builds the library and the tool if needed
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=120)


def test_cli_usage_and_option_errors():
    """exit codes and messages of srgb_to_bt709.m:74-82, 1217-1323"""
    r = _cli()
    assert r.returncode == 1 and r.stdout.startswith("srgb_to_bt709 ?OPTIONS? OUTPUT.y4m\nOPTIONS:\n-frame F.png")
    r = _cli("-gamma", "bogus", "-frame", "a.png", "o.y4m")
    assert r.returncode == 3 and 'option -gamma unknown value "bogus"' in r.stdout
    r = _cli("-fps", "23", "-frame", "a.png", "o.y4m")
    assert r.returncode == 3 and 'option -fps unknown value "23"' in r.stdout
    r = _cli("-bogus", "1", "-frame", "a.png", "o.y4m")
    assert r.returncode == 3 and 'unknown option "-bogus"' in r.stdout
    r = _cli("-gamma", "srgb", "-fps", "30", "o.y4m")
    assert r.returncode == 3 and "either -frame or -frames must be used" in r.stdout
    r = _cli("-frame", "a.png", "-frame", "b.png", "o.y4m")
    assert r.returncode == 3 and "must appear just once" in r.stdout
