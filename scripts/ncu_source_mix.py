#!/usr/bin/env python
"""Dynamic opcode mix + stall samples of one kernel from `ncu --page source --csv`.
usage: ncu -i rep --page source --csv --kernel-name regex:X --launch-count 1 > f.csv; ncu_source_mix.py f.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_inst = tot_samp = 0
for r in data:
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src[:10]
    opc = "IMAD.MOV" if op.startswith("IMAD.MOV") else op.split(".")[0]
    ie = int(r[ix["Instructions Executed"]])
    s = int(r[ix["Warp Stall Sampling (All Samples)"]])
    agg[opc][0] += ie
    agg[opc][1] += s
    agg[opc][2] += 1
    tot_inst += ie
    tot_samp += s
print("total warp-instructions", tot_inst, "stall samples", tot_samp)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{k:12s} static {v[2]:4d} exec {v[0]:10d} {100 * v[0] / tot_inst:5.1f}%  "
          f"samples {v[1]:7d} {100 * v[1] / max(tot_samp, 1):5.1f}%")
for name in hdr:
    if name.startswith("stall_") and "Not Issued" not in name:
        t = sum(int(r[ix[name]] or 0) for r in data)
        if t:
            print(f"{name:26s} {t:8d} {100 * t / max(tot_samp, 1):5.1f}%")
