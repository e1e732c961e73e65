#!/bin/bash
# round 2: pair kernels (128-bit accesses) -- full GPU suite, then same-box A/B
# of the cell kernels through VLCT_PAIR_MASK (bit 0 edge E, 1 face B, 2 update)
mkdir -p gpurun_out
TAG=${TAG:-r2j}
if [ "${TESTS:-1}" = 1 ]; then
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
fi
for rep in 1 2; do
for mask in ${MASKS:-0 7}; do
  for wl in ${WORKLOADS:-ot}; do
  VLCT_PAIR_MASK=$mask timeout 600 python bench.py --workload $wl --steps 6 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/bench_${TAG}_m${mask}_$wl.json 2> gpurun_out/bench_${TAG}_m${mask}_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_m${mask}_$wl.json").read().strip().splitlines()[-1])
    print("mask $mask $wl ms/step %.2f " % d["ms_per_step"], d["clocks"].get("sm_mhz"), {k[2:]: round(v["ms_per_step"], 2) for k,v in d["kernels"].items() if v["ms_per_step"] > 0.1})
except Exception as e:
    print("mask $mask failed", e); print(open("gpurun_out/bench_${TAG}_m${mask}_$wl.err").read()[-1500:])
PY
  done
done
done
