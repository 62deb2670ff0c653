#!/usr/bin/env python
"""Print the key metrics of every kernel in an .ncu-rep (run here, no GPU needed):  python tools/ncu_keys.py file.ncu-rep [out.txt]"""
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from summarize_profiles import KEYS, raw_page  # noqa: E402

rows, units = raw_page(sys.argv[1])
lines = []
for r in rows:
    lines.append("== %s  grid %s block %s" % (r.get("Kernel Name"), r.get("Grid Size"), r.get("Block Size")))
    for k in KEYS:
        if k in r:
            lines.append("  %-80s %s %s" % (k, r[k], units.get(k, "")))
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
