#!/usr/bin/env python
"""Dynamic opcode mix and stall attribution of one kernel from
`ncu -i rep --page source --csv` output. usage: ncu_opmix.py src.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0] != "Address"]


def opof(r):
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    op = m.group(2) if m else "?"
    return "IMAD.MOV" if op.startswith("IMAD.MOV") else op.split(".")[0]


tot = 0
byop = collections.Counter()
samp = collections.Counter()
for r in data:
    ie = int(r[ix["Instructions Executed"]])
    tot += ie
    byop[opof(r)] += ie
    samp[opof(r)] += int(r[ix["# Samples"]])
ts = sum(samp.values())
print("total warp-instructions", tot)
for k, v in byop.most_common(28):
    print(f"{k:12s} {v:11d} {100 * v / tot:5.1f}%  samples {100 * samp[k] / ts:5.1f}%")
for k in ("stall_long_sb", "stall_wait", "stall_math", "stall_short_sb",
          "stall_branch_resolving", "stall_not_selected", "stall_no_inst",
          "stall_dispatch"):
    c = collections.Counter()
    for r in data:
        c[opof(r)] += int(r[ix[k]] or 0)
    t = sum(c.values())
    print(f"{k:24s} {t:7d} ({100 * t / ts:4.1f}%)",
          [(a, round(100 * b / max(t, 1))) for a, b in c.most_common(6)])
