#!/usr/bin/env python
"""Dynamic instruction counts per CUDA-C source line: joins `ncu --page source
--csv` (per-SASS-instruction executed counts) with `nvdisasm -g` line info of
the same cubin. usage: ncu_lines.py src.csv all.dis <mangled-substring> [opfilter]"""
import collections
import csv
import re
import sys

src_csv, dis, sub = sys.argv[1:4]
opf = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]

lines = open(dis).read().split("\n")
start = end = None
for i, l in enumerate(lines):
    if l.startswith(".text.") and sub in l:
        start = i
    elif l.startswith(".text.") and start is not None and i > start:
        end = i
        break
end = end or len(lines)
cur = None
instr = []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        instr.append((int(m.group(1), 16), m.group(3), cur))
if len(data) % len(instr) == 0:
    data = data[:len(instr)]   # ncu repeats the table once per view
assert len(instr) == len(data), (len(instr), len(data))
per = collections.Counter()
perop = collections.defaultdict(collections.Counter)
tot = 0
for (off, op, line), r in zip(instr, data):
    ie = int(r[ix["Instructions Executed"]])
    tot += ie
    if opf and not op.startswith(opf):
        continue
    per[line] += ie
    perop[line][op.split(".")[0] if not op.startswith("IMAD.MOV") else "IMAD.MOV"] += ie
print("total", tot)
for k, v in per.most_common(40):
    ops = ", ".join(f"{o}:{100 * c / v:.0f}%" for o, c in perop[k].most_common(4))
    print(f"{str(k):34s} {v:11d} {100 * v / tot:5.1f}%   {ops}")
