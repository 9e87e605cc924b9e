#!/usr/bin/env python
"""Top SASS instructions of a kernel by warp-stall samples, from an .ncu-rep captured with --set full --import-source on.

    python scripts/ncu_hot.py gpurun_out/prof.ncu-rep [top_n] [context_lines]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None
hdr = None
data = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if data:
            break
        kern = r[1]
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
si = hdr.index("Source")
ci = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
tot = sum(int(r[ci]) for r in data)
print(kern, "total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ci]))[:top]
for i in sorted(order):
    r = data[i]
    stalls = sorted(((int(r[j]), hdr[j]) for j in stall_cols if r[j].isdigit() and int(r[j]) > 0), reverse=True)[:3]
    print(f"{int(r[ci]):7d} {100 * int(r[ci]) / tot:5.1f}%  [{i:5d}] {r[si].strip()[:90]:90s} {stalls}")
    if ctx:
        for k in range(max(0, i - ctx), i):
            print(f"{'':15s} [{k:5d}] {data[k][si].strip()[:90]}")
