"""Helpers for the -m gpu tests: numpy <-> device tensors, a shared context."""
import numpy as np
import torch

import metalbt709decoder_b200 as pkg

_ctx = None


def ctx():
    global _ctx
    if _ctx is None:
        _ctx = pkg.BT709Context(0)
    return _ctx


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def decode(gamma, y, c, alpha=None, w=None):
    """numpy NV12 (h,wy) / (ch,wc) -> numpy (h,w,4) through bt709_decode_nv12"""
    h = y.shape[0]
    w = w if w is not None else y.shape[1]
    ty, tc = dev(y)[None], dev(c)[None]
    ta = dev(alpha)[None] if alpha is not None else None
    out = torch.zeros((1, h, w, 4), dtype=torch.uint8, device="cuda")
    ctx().decode_nv12(gamma, ty, tc, out, alpha_y=ta, w=w, h=h, wait=True)
    return out[0].cpu().numpy()


def parity_report(got, want):
    """(exact fraction, max abs diff) over all channels"""
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    return float((d == 0).mean()), int(d.max())
