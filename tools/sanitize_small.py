#!/usr/bin/env python3
"""Small invocations of every round-2 kernel, for compute-sanitizer (memcheck / racecheck / synccheck):
compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import metalbt709decoder_b200 as pkg
from _oracle import Oracle, bgra_random, frame_random

ora = Oracle()
ctx = pkg.BT709Context(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
ok = True
# k_decode_lin: both rings come from the defaults (opaque 2e, Linear / alpha 3l) + BT709_RING in the environment
for (w, h, F, g, al) in [(256, 8, 3, 3, False), (384, 6, 2, 2, False), (256, 4, 2, 1, True), (4224, 4, 1, 0, False)]:
    fr = [frame_random(w, h, seed=10 + f) for f in range(F)]
    ty = torch.stack([dev(y) for y, _ in fr]); tc = torch.stack([dev(c) for _, c in fr])
    an = [frame_random(w, h, seed=20 + f, full_range=True)[0] for f in range(F)] if al else None
    ta = torch.stack([dev(a) for a in an]) if al else None
    out = torch.zeros((F, h, w, 4), dtype=torch.uint8, device="cuda")
    ctx.decode_nv12(g, ty, tc, out, alpha_y=ta, wait=True)
    got = out.cpu().numpy()
    for f in range(F):
        ok &= bool(np.array_equal(got[f], ora.decode_nv12(g, fr[f][0], fr[f][1], alpha_y=an[f] if al else None)))
# k_rescale_pair: opaque, translucent (redo path)
for alpha in (255, None):
    src = bgra_random(192, 80, seed=3, alpha=alpha)
    out = torch.zeros((1, 107, 256, 4), dtype=torch.uint8, device="cuda")
    ctx.rescale_bgra(dev(src)[None], out, wait=True)
    ok &= bool(np.array_equal(out[0].cpu().numpy(), ora.rescale_bgra(src, 256, 107)))
# k_fused
for g in (3, 2):
    y, c = frame_random(256, 72, seed=5)
    out = torch.zeros((2, 96, 342, 4), dtype=torch.uint8, device="cuda")
    ctx.decode_rescale_nv12_fused(g, dev(y)[None].repeat(2, 1, 1), dev(c)[None].repeat(2, 1, 1),
                                  torch.zeros((2, 72, 256, 4), dtype=torch.uint8, device="cuda"), out, wait=True)
    want = ora.rescale_bgra(ora.decode_nv12(g, y, c), 342, 96)
    ok &= bool(np.array_equal(out[0].cpu().numpy(), want)) and bool(np.array_equal(out[1].cpu().numpy(), want))
# encode (split launcher path untouched here), packed 4:4:4 with a curve argument
src = bgra_random(256, 16, seed=7)
ey = torch.zeros((1, 16, 256), dtype=torch.uint8, device="cuda"); ec = torch.zeros((1, 8, 256), dtype=torch.uint8, device="cuda")
ctx.encode_bgra_to_nv12(1, 0, dev(src)[None], ey, ec, wait=True)
wy, wc = ora.encode_bgra_nv12(1, 0, src)
ok &= bool(np.array_equal(ey[0].cpu().numpy(), wy) and np.array_equal(ec[0].cpu().numpy(), wc))
words = torch.arange(64 * 64, dtype=torch.int32, device="cuda").view(64, 64) * 4099
o1 = torch.zeros_like(words); o2 = torch.zeros_like(words)
ctx.convert_packed444(words, o1, gamma=3); ctx.unconvert_packed444(o1, o2, gamma=3, wait=True)
print("sanitize_small:", "ok" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
