#!/usr/bin/env python3
"""One process, two devices: a context per device, same frames decoded/rescaled/encoded on both, results equal.
Run on a box with >= 2 GPUs: python tools/two_devices_check.py"""
import sys
import torch
sys.path.insert(0, ".")
import metalbt709decoder_b200 as pkg

assert torch.cuda.device_count() >= 2, "needs two GPUs"
w, h, F = 1920, 1080, 4
g = torch.Generator().manual_seed(7)
y = torch.randint(16, 236, (F, h, w), dtype=torch.uint8, generator=g)
c = torch.randint(16, 241, (F, h // 2, w), dtype=torch.uint8, generator=g)
outs = []
for dev in (0, 1):
    ctx = pkg.BT709Context(dev)
    with torch.cuda.device(dev):
        ty, tc = y.cuda(dev), c.cuda(dev)
        o = torch.zeros((F, h, w, 4), dtype=torch.uint8, device=f"cuda:{dev}")
        ctx.decode_nv12(3, ty, tc, o, wait=True)
        big = torch.zeros((F, 1440, 2560, 4), dtype=torch.uint8, device=f"cuda:{dev}")
        ctx.rescale_bgra(o, big, wait=True)
        ey = torch.zeros((F, h, w), dtype=torch.uint8, device=f"cuda:{dev}")
        ec = torch.zeros((F, h // 2, w), dtype=torch.uint8, device=f"cuda:{dev}")
        ctx.encode_bgra_to_nv12(1, 0, o, ey, ec, wait=True)
        outs.append((o.cpu(), big.cpu(), ey.cpu(), ec.cpu()))
for a, b in zip(*outs):
    assert torch.equal(a, b)
print("two devices in one process: identical results", [tuple(t.shape) for t in outs[0]])
