#!/usr/bin/env python
"""Hot regions of a kernel from an .ncu-rep (source page, SASS view): consecutive instructions with PC samples, with the dominant stall reason.
usage: tools/ncu_sass_hot.py rep.ncu-rep [min_samples]"""
import csv, subprocess, sys
rep = sys.argv[1]; mins = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
h = rows[hi]; idx = {k: i for i, k in enumerate(h)}
stall = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
data = [r for r in rows[hi + 1:] if len(r) > idx["# Samples"] and r[idx["# Samples"]].isdigit()]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("total samples", tot)
for n, r in enumerate(data):
    s = int(r[idx["# Samples"]])
    if s >= mins:
        st = sorted(((int(r[idx[k]] or 0), k[6:]) for k in stall), reverse=True)[:3]
        print(f"{n:5d} {s:5d} {100*s/tot:5.1f}%  {r[idx['Source']].strip()[:70]:70s} {st}")
