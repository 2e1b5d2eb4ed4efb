#!/usr/bin/env python
"""Executed-instruction mix of one captured kernel: `ncu -i X.ncu-rep --page source --csv --print-source sass` folded by opcode and
normalised per ray.   usage: tools/ncu_opmix.py <capture.ncu-rep> <rays> [out.txt]"""
import collections
import csv
import re
import subprocess
import sys

rep, rays = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
hdr = rows[1]
ia, ie, it, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
byop, samp, tot, thr, static = collections.Counter(), collections.Counter(), 0, 0, 0
for r in rows[2:]:
    if len(r) <= it:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia].strip())
    op = m.group(2) if m else r[ia].strip()[:10]
    ex = int(r[ie] or 0)
    byop[op] += ex
    samp[op] += int(r[ism] or 0)
    tot += ex
    thr += int(r[it] or 0)
    static += 1
lines = [f"# {kernel}", f"# {static} SASS instructions; executed: {tot} warp-instructions = {tot / rays:.1f} per ray, {thr / max(tot, 1):.2f} active lanes per instruction",
         f"# {'opcode':10s} {'warp-instr/ray':>15s} {'share %':>8s} {'stall samples':>14s}"]
for op, c in byop.most_common(32):
    lines.append(f"{op:12s} {c / rays:15.2f} {100 * c / tot:8.1f} {samp[op]:14d}")
out = "\n".join(lines) + "\n"
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(out)
print(out)
