#!/usr/bin/env python3
"""The driver's HBM peak (MEASURED_PEAKS.json: torch b.copy_(a) over 1 Gi bf16 elements, best of 10) as a burst and
repeated for ~2 s, with SM / memory clocks and power sampled meanwhile: what the copy figure itself does under the
1000 W cap.  python tools/copy_sustained.py"""
import json
import sys
import time
import torch
sys.path.insert(0, ".")
from bench import ClockSampler

a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)
nbytes = 2 * a.numel() * 2


def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = ClockSampler(0, lms=25)
    s.start()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    c = s.stop()
    return nbytes * n / (e0.elapsed_time(e1) / 1e3) / 1e9, c


for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
time.sleep(0.5)
best = 0.0
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9)
time.sleep(0.5)
burst, cb = run(20)
sustained, cs = run(3000)
print(json.dumps({"copy_best_of_10_GBps": best, "copy_20_GBps": burst, "clocks_20": cb, "copy_3000_GBps": sustained,
                  "clocks_3000": cs}))
