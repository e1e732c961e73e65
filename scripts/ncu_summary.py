#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per launch the metrics that the
roofline / bound analysis needs. usage: ncu_summary.py file.ncu-rep [more...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum",
    "smsp__inst_executed.sum",
    "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__block_size", "launch__grid_size",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"== {path}")
        stall_cols = [i for i, h in enumerate(hdr)
                      if h.startswith("smsp__average_warps_issue_stalled")
                      and h.endswith("_per_issue_active.ratio")]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"-- {d['Kernel Name'][:70]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
            for k in KEYS:
                if k in d:
                    print(f"   {k:68s} {d[k]:>18s} {units[hdr.index(k)]}")
            stalls = sorted(((float(r[i].replace(',', '') or 0), hdr[i]) for i in stall_cols),
                            reverse=True)[:6]
            for v, h in stalls:
                print(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.2f}")


if __name__ == "__main__":
    main()
