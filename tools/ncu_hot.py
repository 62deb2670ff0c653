#!/usr/bin/env python
"""Stall samples by SASS region of a captured kernel:  python tools/ncu_hot.py file.ncu-rep [n_regions]
Splits the instruction stream at the USETMAXREG / BAR / SYNCS boundaries and prints samples + instruction counts per
region and the top instructions."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ia, isrc, isamp, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = []
for r in rows[2:]:
    if len(r) <= iexe:
        continue
    try:
        ins.append((r[isrc].strip(), int(r[isamp] or 0), int(r[iexe] or 0)))
    except ValueError:
        pass
tot = sum(s for _, s, _ in ins) or 1
tote = sum(e for _, _, e in ins) or 1
print("instructions %d, samples %d, executed %d" % (len(ins), tot, tote))
# regions split at barriers
reg, start = [], 0
for i, (src, s, e) in enumerate(ins):
    if src.startswith(("BAR", "USETMAXREG", "WARPSYNC")) or "SYNCS.PHASECHK" in src or "SYNCS.ARRIVE" in src:
        reg.append((start, i))
        start = i
reg.append((start, len(ins)))
print("-- regions (first idx, n instr, %samples, %executed, first instruction)")
for a, b in reg:
    s = sum(x[1] for x in ins[a:b])
    e = sum(x[2] for x in ins[a:b])
    if s * 100.0 / tot >= 1.0 or e * 100.0 / tote >= 1.0:
        print("%5d %5d  %5.1f%%  %5.1f%%  %s" % (a, b - a, 100.0 * s / tot, 100.0 * e / tote, ins[a][0][:60]))
print("-- top instructions")
for i in sorted(range(len(ins)), key=lambda k: -ins[k][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%5d  %5.2f%%  %s" % (i, 100.0 * ins[i][1] / tot, ins[i][0][:90]))
