#!/usr/bin/env python
"""Per-launch table of an `ncu --csv --metrics ...` log of one step
(scripts/gpu_r2a.sh): time, fp64 / all warp-instructions, DRAM bytes, and the
per-32-cell-update totals bench.py's roofline_fp64 uses.
usage: step_metrics.py log.csv [cells]"""
import csv
import sys


def main():
    path = sys.argv[1]
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 512 ** 3
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    agg = {}
    for r in data:
        d = dict(zip(hdr, r))
        key = (int(d["ID"]), d["Kernel Name"][:64])
        agg.setdefault(key, {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    tot = {}
    print(f"{'id':>3} {'kernel':64s} {'ms':>8} {'fp64 Ginst':>10} {'all Ginst':>10} {'rd GB':>7} {'wr GB':>7}")
    for (i, k), m in sorted(agg.items()):
        print(f"{i:3d} {k:64s} {m.get('gpu__time_duration.sum', 0) / 1e6:8.3f} "
              f"{m.get('smsp__inst_executed_pipe_fp64.sum', 0) / 1e9:10.3f} "
              f"{m.get('smsp__inst_executed.sum', 0) / 1e9:10.3f} "
              f"{m.get('dram__bytes_read.sum', 0) / 1e9:7.2f} {m.get('dram__bytes_write.sum', 0) / 1e9:7.2f}")
        for kk, vv in m.items():
            tot[kk] = tot.get(kk, 0) + vv
    w = cells / 32
    print(f"step: {tot['gpu__time_duration.sum'] / 1e6:.2f} ms (ncu: serialised, cold, short -> not power-capped), "
          f"fp64 warp-inst per 32 cell-updates {tot['smsp__inst_executed_pipe_fp64.sum'] / w:.1f}, "
          f"all warp-inst per 32 cell-updates {tot['smsp__inst_executed.sum'] / w:.1f}, "
          f"DRAM {(tot['dram__bytes_read.sum'] + tot['dram__bytes_write.sum']) / 1e9:.1f} GB")


if __name__ == "__main__":
    main()
