#!/usr/bin/env python
"""SASS instruction-count summary of the built library, per kernel (runs without a GPU: cuobjdump -sass).
Shows that the hot kernels are Blackwell-native - UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk),
UCGABAR_ARV / _WAIT (barrier.cluster) - and what they still pay in local-memory spills (STL / LDL) and FP32 FMA.
usage: python tools/sass_summary.py [path/to/lib.so] > profiles/sass_rNN.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rankaae_b200", "librankaae_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for l in res.splitlines():
    m = re.search(r"Function (\S+):", l)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in l:
        usage[cur] = l.strip()
        cur = None
kernels = OrderedDict()
cur = None
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        kernels[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", l)
    if m and cur:
        kernels[cur][m.group(1)] += 1
KEYS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "CCTL", "LDGSTS", "FFMA", "MUFU", "LDS", "STS",
        "LDG", "STG", "LDL", "STL", "BAR", "SHFL"]
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: instruction counts per kernel (static), sm_100a")
print(f"# {'kernel':40s} {'total':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
for name, c in kernels.items():
    tot = sum(c.values())
    if tot < 50:
        continue
    print(f"{demangle(name):42s} {tot:7d} " + " ".join(f"{c.get(k, 0):7d}" for k in KEYS))
print()
for name in kernels:
    if name in usage and sum(kernels[name].values()) >= 50:
        print(f"{demangle(name):42s} {usage[name]}")
