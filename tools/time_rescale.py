#!/usr/bin/env python3
"""Time bt709_rescale_bgra for arbitrary sizes: python tools/time_rescale.py IW IH OW OH [frames] [noise|smooth] [opaque]
(opaque: A = 255 everywhere, what pass 1 hands over for video without an alpha stream)"""
import sys
import torch
sys.path.insert(0, ".")
import metalbt709decoder_b200 as pkg

iw, ih, ow, oh = map(int, sys.argv[1:5])
frames = int(sys.argv[5]) if len(sys.argv) > 5 else 64
smooth = len(sys.argv) > 6 and sys.argv[6] != "noise"  # smooth gradients + 2 LSB of noise instead of white noise
opaque = len(sys.argv) > 7
ctx = pkg.BT709Context(0)
src = torch.randint(0, 256, (frames, ih, iw, 4), dtype=torch.uint8, device="cuda")
if smooth:
    yy = torch.arange(ih, device="cuda").view(1, ih, 1, 1)
    xx = torch.arange(iw, device="cuda").view(1, 1, iw, 1)
    ch = torch.arange(4, device="cuda").view(1, 1, 1, 4)
    ff = torch.arange(frames, device="cuda").view(frames, 1, 1, 1)
    base = (xx * (ch + 1) // 8 + yy * (4 - ch) // 6 + ff * 3) % 256
    src = ((base + (src & 3)) % 256).to(torch.uint8)
if opaque:
    src[..., 3] = 255
dst = torch.zeros((frames, oh, ow, 4), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ctx.rescale_bgra(src, dst)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    ctx.rescale_bgra(src, dst)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
gb = frames * (iw * ih + ow * oh) * 4 / 1e9
print(f"{'smooth' if smooth else 'noise'}{' opaque' if opaque else ''} {iw}x{ih} -> {ow}x{oh} x{frames}: {ms:.3f} ms  {frames*iw*ih/ms/1e6:.1f} Gpix/s in  {frames*ow*oh/ms/1e6:.1f} Gpix/s out  {gb/ms*1e3:.0f} GB/s")
