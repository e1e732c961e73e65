#!/bin/bash
# same-box A/B of the e2e (HOST, pipelined) path over VLCT_PAIR_MASK values
mkdir -p gpurun_out
TAG=${TAG:-e2e}
for rep in 1 2; do
for mask in ${MASKS:-14 30 6}; do
  VLCT_PAIR_MASK=$mask timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_${TAG}_m$mask.json 2> gpurun_out/bench_${TAG}_m$mask.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_m$mask.json").read().strip().splitlines()[-1])
    e=d["e2e"]
    print("mask $mask dev %.2f ms  e2e %.1f ms (h2d %.1f d2h %.1f GB/s) two_calls %.1f" % (d["ms_per_step"], e["ms_per_step"], e["pcie_gbs_h2d"], e["pcie_gbs_d2h"], e["two_calls"]["ms_per_step"]))
except Exception as ex:
    print("mask $mask failed", ex); print(open("gpurun_out/bench_${TAG}_m$mask.err").read()[-1500:])
PY
done
done
