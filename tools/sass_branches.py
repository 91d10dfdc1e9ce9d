#!/usr/bin/env python
"""Branch diamonds (BSSY) and slow-path calls (CALL) per CUDA source line and per stage function of raae_train_kernel.

The fused kernel runs two warps per scheduler, so independent work inside an unrolled loop has to be interleaved by ptxas;
a branch diamond per element / per row group (a `cond ? cheap : expensive` the compiler turns into a branch, an IEEE division
or square root with its slow-path call, an `if (row < nv)` around a body with stores) cuts the loop into scheduling regions
and serialises it.  This tool is how the round-2 diamonds were found (dropout-mask source, Softplus threshold, AdamW element
update, loss row pass, Box-Muller).  Needs build/rankaae_b200.o (python __graft_entry__.py --force), cuobjdump, nvdisasm.
usage: python tools/sass_branches.py [top_n]"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOP = int(sys.argv[1]) if len(sys.argv) > 1 else 25
obj = os.path.join(ROOT, "build", "rankaae_b200.o")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
lines = text.split("\n")
start = next(i for i, l in enumerate(lines) if ".text._ZN4raae17raae_train_kernel" in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//---") and ".text." in lines[i]), len(lines))
lines = lines[start:end]
srcs = {}


def src(f, n):
    if f not in srcs:
        try:
            srcs[f] = open(os.path.join(ROOT, "rankaae_b200", "csrc", f)).read().split("\n")
        except OSError:
            srcs[f] = []
    return srcs[f][n - 1].strip()[:100] if 0 < n <= len(srcs[f]) else ""


func, cur = "raae_train_kernel (body)", None
per_line, per_func, n_instr = Counter(), Counter(), Counter()
for l in lines:
    m = re.match(r"^\$_ZN4raae17raae_train_kernel\w+\$_ZN4raae\d+(\w+?)E", l.strip())
    if m and l.strip().endswith(":"):
        func = re.sub(r"(ERK|ILb).*", "", m.group(1))
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?(\w[\w.]*)", l)
    if m and cur:
        op = m.group(2).split(".")[0]
        n_instr[func] += 1
        if op in ("BSSY", "CALL"):
            per_line[(func, cur, op)] += 1
            per_func[(func, op)] += 1
print("# per stage function: SASS instructions, branch diamonds (BSSY), calls (CALL: nested stage calls + slow paths)")
for f, n in sorted(n_instr.items(), key=lambda kv: -kv[1]):
    print(f"{f:28s} {n:7d} instr  {per_func[(f, 'BSSY')]:4d} BSSY  {per_func[(f, 'CALL')]:4d} CALL")
print(f"\n# top {TOP} source lines by BSSY + CALL")
for (f, (sf, ln), op), c in per_line.most_common(TOP):
    print(f"{c:4d} {op:5s} {f:22s} {sf}:{ln}  {src(sf, ln)}")
