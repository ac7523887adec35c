#!/usr/bin/env python3
"""Time bt709_decode_nv12 / bt709_encode_bgra_to_nv12 for arbitrary sizes:
python tools/time_decode.py W H [frames] [pad_bytes] [encode [gin gout]]"""
import sys
import torch
sys.path.insert(0, ".")
import metalbt709decoder_b200 as pkg

w, h = int(sys.argv[1]), int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 64
pad = int(sys.argv[4]) if len(sys.argv) > 4 else 0
encode = len(sys.argv) > 5
gin = int(sys.argv[6]) if len(sys.argv) > 6 else 1
gout = int(sys.argv[7]) if len(sys.argv) > 7 else 0
ctx = pkg.BT709Context(0)
cw = (w + 1) // 2 * 2
y = torch.randint(16, 236, (frames, h, w + pad), dtype=torch.uint8, device="cuda")[:, :, :w]
c = torch.randint(16, 241, (frames, (h + 1) // 2, cw + pad), dtype=torch.uint8, device="cuda")[:, :, :cw]
# (encode reads `out`: random BGRA, A = 255 - all-zero input runs 3-5 % faster than real content, the table reads collapse)
out = (torch.randint(0, 256, (frames, h, w + pad, 4), dtype=torch.uint8, device="cuda") if encode
       else torch.zeros((frames, h, w + pad, 4), dtype=torch.uint8, device="cuda"))[:, :, :w]
def run():
    if encode:
        ctx.encode_bgra_to_nv12(gin, gout, out, y, c)
    else:
        ctx.decode_nv12(3, y, c, out)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{'encode %d->%d' % (gin, gout) if encode else 'decode'} {w}x{h} x{frames} pad {pad}: {ms:.3f} ms  {frames*w*h/ms/1e6:.1f} Gpix/s  {frames*w*h*5.5/ms/1e6:.0f} GB/s")
