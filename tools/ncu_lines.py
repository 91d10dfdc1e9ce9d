#!/usr/bin/env python
"""Summarises an .ncu-rep per CUDA source line: samples, instructions, stall breakdown.
usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = None, None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        lines.append((cur_file, int(r[0]), r[1].strip(), d))
def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot_s = sum(I(d["# Samples"]) for _, _, _, d in lines)
tot_i = sum(I(d["Instructions Executed"]) for _, _, _, d in lines)
print(f"total samples {tot_s}  total warp-instructions {tot_i}")
byfile = defaultdict(lambda: [0, 0])
for f, ln, src, d in lines:
    byfile[f][0] += I(d["# Samples"]); byfile[f][1] += I(d["Instructions Executed"])
for f, (s, i) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:24s} samples {100*s/tot_s:5.1f}%  instr {100*i/tot_i:5.1f}%")
stalls = [k for k in lines[0][3] if k.startswith("stall_")]
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s}  top stalls | source")
for f, ln, src, d in sorted(lines, key=lambda x: -I(x[3]["# Samples"]))[:top]:
    s = I(d["# Samples"])
    st = sorted(((int(d[k]), k[6:]) for k in stalls if d[k].isdigit()), reverse=True)[:3]
    sts = " ".join(f"{k}:{100*v/max(s,1):.0f}%" for v, k in st)
    print(f"{f+':'+str(ln):28s} {100*s/tot_s:6.2f} {100*I(d['Instructions Executed'])/tot_i:6.2f}  {sts:40s} | {src[:90]}")
