#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass), as markdown on stdout.
Highlights the mnemonics that prove the Blackwell paths: UTCHMMA / UTCBAR (tcgen05.mma / commit), LDTM / STTM (tcgen05.ld/st),
UBLKCP (cp.async.bulk), UTMALDG (TMA tensor load), LDGSTS (cp.async), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (packed fp32).

    python tools/sass_histogram.py > profiles/sass_r2.md"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "implicit_depth_b200", "csrc", "liblidf_query.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "HMMA",
       "MUFU", "SHFL", "ATOM", "RED", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("# SASS opcode histogram per kernel (`cuobjdump -sass implicit_depth_b200/csrc/liblidf_query.so`, sm_100a)\n")
print("| kernel | instructions | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
tot = collections.Counter()
for k, h in sorted(hist.items(), key=lambda kv: -sum(kv[1].values())):
    agg = {key: sum(v for op, v in h.items() if op == key or (key in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "ATOM", "RED", "SHFL", "MUFU", "SYNCS", "HMMA") and op.startswith(key))) for key in KEY}
    print(f"| {k[:48]} | {sum(h.values())} | " + " | ".join(str(agg[key]) if agg[key] else "" for key in KEY) + " |")
    tot.update(agg)
print("| **all kernels** | " + str(sum(sum(h.values()) for h in hist.values())) + " | " + " | ".join(str(tot[key]) for key in KEY) + " |")
