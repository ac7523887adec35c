#!/usr/bin/env python3
"""Kernel time of SMALL decode calls (a few frames per call), free of Python / launch overhead: the calls are
captured into one CUDA graph (N launches back to back on one stream) and the graph is timed with CUDA events.
python tools/time_small.py W H frames [launches_per_graph] [gamma] [decode|rescale|twopass|fused|encode]
(rescale / twopass / fused scale by 4/3 like the 1080p -> 1440p configuration)"""
import sys
import torch
sys.path.insert(0, ".")
import metalbt709decoder_b200 as pkg

w, h, frames = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 200
gamma = int(sys.argv[5]) if len(sys.argv) > 5 else 3
mode = sys.argv[6] if len(sys.argv) > 6 else "decode"
ow, oh = w * 4 // 3, h * 4 // 3
ctx = pkg.BT709Context(0)
# a ring of independent inputs/outputs far larger than L2, so every launch reads cold data
ring = max(2, min(64, (1 << 30) // (frames * w * h * 6)))
ys = [torch.randint(16, 236, (frames, h, w), dtype=torch.uint8, device="cuda") for _ in range(ring)]
cs = [torch.randint(16, 241, (frames, h // 2, w), dtype=torch.uint8, device="cuda") for _ in range(ring)]
outs = [torch.empty((frames, h, w, 4), dtype=torch.uint8, device="cuda") for _ in range(ring)]
bigs = [torch.empty((frames, oh, ow, 4), dtype=torch.uint8, device="cuda") for _ in range(ring)] if mode in ("rescale", "twopass", "fused") else None
if mode in ("rescale", "encode"):
    for o in outs:
        o.copy_(torch.randint(0, 256, o.shape, dtype=torch.uint8, device="cuda"))
        o[..., 3] = 255


def call(i, st):
    if mode == "decode":
        ctx.decode_nv12(gamma, ys[i], cs[i], outs[i], stream=st)
    elif mode == "rescale":
        ctx.rescale_bgra(outs[i], bigs[i], stream=st)
    elif mode == "twopass":
        ctx.decode_rescale_nv12(gamma, ys[i], cs[i], outs[i], bigs[i], stream=st)
    elif mode == "fused":
        ctx.decode_rescale_nv12_fused(gamma, ys[i], cs[i], outs[i], bigs[i], stream=st)
    elif mode == "encode":
        ctx.encode_bgra_to_nv12(1, 0, outs[i], ys[i], cs[i], stream=st)
    else:
        raise SystemExit("mode?")


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(ring):
        call(i, s)
s.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    for i in range(n):
        call(i % ring, torch.cuda.current_stream())
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / n
print(f"{mode} {w}x{h} x{frames} gamma {gamma}: {us:.2f} us per launch ({n} launches in one graph, ring of {ring})  "
      f"{frames*w*h/us/1e3:.1f} Gpix/s")
