#!/usr/bin/env python3
"""List the loops (backward branches) of one kernel in a cubin / shared library with their length and opcode mix:
python tools/sass_loops.py FILE KERNEL_SUBSTRING [min_len]   (static counts: no GPU needed)"""
import collections
import re
import subprocess
import sys

path, key = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 8
txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
ins, cur = [], None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or key not in cur:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f"{key}: {len(ins)} instructions")
addr_index = {a: i for i, (a, _) in enumerate(ins)}
def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr_index:
        body = ins[addr_index[tgt]: i + 1]
        if len(body) < min_len:
            continue
        h = collections.Counter(opcode(x) for _, x in body)
        print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr  " + " ".join(f"{k}:{v}" for k, v in h.most_common(14)))
