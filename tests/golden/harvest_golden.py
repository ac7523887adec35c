#!/usr/bin/env python3
"""Harvest the reference's own known-answer vectors into tests/golden/reference_vectors.json.

Run in the build container only (needs /root/reference); the JSON is committed so
the GPU box never needs the reference tree.

    python tests/golden/harvest_golden.py

It parses the XCTest sources under /root/reference/EmptyiOSTests — it does not
execute them (they need an Apple device).  For every live test method it records
  * the integer inputs the test assigns (Rin, R1..B4, ...),
  * every `{ int v = NAME; int expectedVal = EXPR; XCTAssert(v == expectedVal) }`
    block, with EXPR evaluated against those inputs (e.g. `Rin - 1`),
  * which reference functions / back-end enums / gamma constants the body names,
so that tests/test_oracle_golden.py can replay each one through the oracle.
Block comments (`/* */`, e.g. the dead first half of AppleEncodeDecodeBT709Tests.m
and all of MetalSRGBDecoderTests.m) and `//` lines are stripped first, so only
live expectations are harvested.
"""
import json
import os
import re
import sys

REF = "/root/reference/EmptyiOSTests"
FILES = [
    "MetalBT709DecoderTests.m",
    "AppleEncodeDecodeBT709Tests.m",
    "CoreImageMetalFilterTests.m",
]
CALLS = [
    "BT709_convertLinearRGBToYCbCr", "BT709_convertYCbCrToRGB",
    "BT709_from_sRGB_convertRGBToYCbCr", "BT709_to_sRGB_convertYCbCrToRGB",
    "Apple196_from_sRGB_convertRGBToYCbCr", "Apple196_to_sRGB_convertYCbCrToRGB",
    "sRGB_from_sRGB_convertRGBToYCbCr", "sRGB_to_sRGB_convertYCbCrToRGB",
    "BT709_average_pixel_values", "sRGB_from_sRGB_average_cbcr",
    "convert_srgb_to_bt709", "convert_bt709_to_srgb",
]
ENUMS = [
    "BGRAToBT709ConverterSoftware", "BGRAToBT709ConverterVImage", "BGRAToBT709ConverterMetal",
    "BT709GammaSrgb", "BT709GammaApple", "BT709GammaLinear",
]


def strip_comments(src):
    """Remove /* */ and // comments but keep newlines so line numbers survive."""
    out = []
    i, n = 0, len(src)
    while i < n:
        if src.startswith("/*", i):
            j = src.find("*/", i + 2)
            j = n if j < 0 else j + 2
            out.append(re.sub(r"[^\n]", " ", src[i:j]))
            i = j
        elif src.startswith("//", i):
            j = src.find("\n", i)
            j = n if j < 0 else j
            i = j
        elif src[i] == '"':  # string literal (format strings hold %3d etc.)
            j = i + 1
            while j < n and src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            out.append(src[i:j + 1])
            i = j + 1
        else:
            out.append(src[i])
            i += 1
    return "".join(out)


def eval_expr(expr, env):
    expr = expr.strip()
    m = re.fullmatch(r"(\w+)\s*([+-])\s*(\d+)", expr)
    if m:
        base = eval_expr(m.group(1), env)
        if base is None:
            return None
        return base + int(m.group(3)) if m.group(2) == "+" else base - int(m.group(3))
    if re.fullmatch(r"-?\d+", expr):
        return int(expr)
    if re.fullmatch(r"0x[0-9a-fA-F]+", expr):
        return int(expr, 16)
    return env.get(expr)


def parse_method(name, body, first_line):
    env = {}
    order = []
    checks = []
    # walk statements in order so `Gin = Rin;` sees the value of Rin
    pos = 0
    pat_assign = re.compile(r"(?:\b(?:const\s+)?(?:int|uint32_t)\s+)?\b([A-Za-z_]\w*)\s*=\s*([^;=]+);")
    pat_check = re.compile(r"int\s+v\s*=\s*(\w+)\s*;\s*int\s+expectedVal\s*=\s*([^;]+);")
    events = []
    for m in pat_check.finditer(body):
        events.append((m.start(), "check", m))
    check_spans = [(m.start(), m.end()) for _, _, m in events]
    for m in pat_assign.finditer(body):
        if any(a <= m.start() < b for a, b in check_spans):
            continue
        events.append((m.start(), "assign", m))
    events.sort(key=lambda e: e[0])
    for _, kind, m in events:
        if kind == "assign":
            var, expr = m.group(1), m.group(2)
            if var in ("v", "expectedVal", "result", "worked"):
                continue
            val = eval_expr(expr, env)
            if val is not None:
                env[var] = val
                if var not in order:
                    order.append(var)
        else:
            val = eval_expr(m.group(2), env)
            checks.append({"var": m.group(1), "expected": val, "expr": m.group(2).strip()})
    calls = [c for c in CALLS if re.search(r"\b" + c + r"\b", body)]
    enums = {}
    for e in ENUMS:
        hits = [mm.start() for mm in re.finditer(r"\b" + e + r"\b", body)]
        if hits:
            enums[e] = len(hits)
    # which back-end each of encodeType/decodeType/type is bound to (live assignment only)
    binds = {}
    for mm in re.finditer(r"BGRAToBT709ConverterTypeEnum\s+(\w+)\s*=\s*(\w+)\s*;", body):
        binds[mm.group(1)] = mm.group(2)
    gam = {}
    for mm in re.finditer(r"BT709Gamma\s+(\w+)\s*=\s*(\w+)\s*;", body):
        gam[mm.group(1)] = mm.group(2)
    return {
        "name": name, "line": first_line, "inputs": {k: env[k] for k in order},
        "checks": checks, "calls": calls, "backends": binds, "gammas": gam,
    }


def harvest_file(fname):
    raw = open(os.path.join(REF, fname), encoding="utf-8", errors="replace").read()
    src = strip_comments(raw)
    tests = []
    heads = list(re.finditer(r"^- \(void\)\s*(test\w+)\s*\{", src, re.M))
    for i, h in enumerate(heads):
        end = heads[i + 1].start() if i + 1 < len(heads) else len(src)
        # a method ends before the next "- (" helper as well
        nxt = re.search(r"^- \(", src[h.end():end], re.M)
        if nxt:
            end = h.end() + nxt.start()
        body = src[h.end():end]
        line = src.count("\n", 0, h.start()) + 1
        tests.append(parse_method(h.group(1), body, line))
    return tests


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; golden vectors are already committed")
    out = {"_source": "harvested from /root/reference/EmptyiOSTests by tests/golden/harvest_golden.py",
           "files": {}}
    for f in FILES:
        out["files"][f] = harvest_file(f)
    # pinned histogram counts live in comments next to the asserts; keep them explicit here
    out["histograms"] = {
        # CoreImageMetalFilterTests.m:534,541-551 / :671,679-689 / :810,826-829
        "bt709": {"gamma": 3, "exact": 2753221, "buckets": [2753221, 12653553, 1006705, 295446, 67339, 952]},
        "apple": {"gamma": 0, "exact": 2753405, "buckets": [2753405, 13499212, 524599]},
        "srgb": {"gamma": 1, "exact": 2753772, "buckets": [2753772, 13893861, 129583]},
    }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    for f, ts in out["files"].items():
        print(f, len(ts), "tests,", sum(len(t["checks"]) for t in ts), "checks")


if __name__ == "__main__":
    main()
