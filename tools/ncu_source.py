#!/usr/bin/env python
"""Top SASS instructions of an .ncu-rep by stall samples / executed count.  Usage: ncu_source.py rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot_s = sum(int(r[idx["# Samples"]]) for r in body)
tot_i = sum(int(r[idx["Instructions Executed"]]) for r in body)
print("total samples", tot_s, "total warp-inst", tot_i)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("--- by samples")
for r in sorted(body, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
    top = sorted(((int(r[idx[s]]), s) for s in stalls), reverse=True)[:2]
    print(f"{int(r[idx['# Samples']]):7d} {100*int(r[idx['# Samples']])/tot_s:5.1f}%  ex={int(r[idx['Instructions Executed']]):9d}  {body.index(r):4d} {r[idx['Source']].strip():60s} {top}")
print("--- by executed")
for r in sorted(body, key=lambda r: -int(r[idx["Instructions Executed"]]))[:n // 2]:
    print(f"{int(r[idx['Instructions Executed']]):9d} {100*int(r[idx['Instructions Executed']])/tot_i:5.1f}%  {body.index(r):4d} {r[idx['Source']].strip()}")
