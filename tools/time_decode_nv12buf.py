#!/usr/bin/env python3
"""Time bt709_decode_nv12 on the layout a decoder hands over: ONE allocation per frame, the CbCr plane right behind the Y
plane (frame stride of both planes = 1.5 * h * pitch, so the batch is not one tall image per plane):
python tools/time_decode_nv12buf.py W H [frames] [gamma]"""
import sys
import torch
sys.path.insert(0, ".")
import metalbt709decoder_b200 as pkg

w, h = int(sys.argv[1]), int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 64
gamma = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = pkg.BT709Context(0)
buf = torch.randint(16, 236, (frames, h + h // 2, w), dtype=torch.uint8, device="cuda")
y, c = buf[:, :h, :], buf[:, h:, :]
out = torch.zeros((frames, h, w, 4), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ctx.decode_nv12(gamma, y, c, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    ctx.decode_nv12(gamma, y, c, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"decode (Y|CbCr per frame) {w}x{h} x{frames} gamma {gamma}: {ms:.3f} ms  {frames*w*h/ms/1e6:.1f} Gpix/s  {frames*w*h*5.5/ms/1e6:.0f} GB/s")
