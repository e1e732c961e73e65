#!/bin/bash
# round 2: fused CT march -- GPU suite, then same-box A/B: separate kernels
# (pair_kernels 2) vs the march (10) in its build variants
mkdir -p gpurun_out
TAG=${TAG:-r2k}
if [ "${TESTS:-1}" = 1 ]; then
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
fi
run() {  # name mask lib
  VLCT_PAIR_MASK=$2 VLCT_B200_LIB=$PWD/$3 timeout 600 python bench.py --workload ${WL:-ot} --steps 6 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/bench_${TAG}_$1.json 2> gpurun_out/bench_${TAG}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1 ms/step %.2f " % d["ms_per_step"], d["clocks"].get("sm_mhz"), {k[2:]: round(v["ms_per_step"], 2) for k,v in d["kernels"].items() if v["ms_per_step"] > 0.1})
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/bench_${TAG}_$1.err").read()[-1500:])
PY
}
for rep in 1 2; do
run sep 2 enzo-e_b200/csrc/libvlct_b200.so
run ct4 10 enzo-e_b200/csrc/libvlct_b200.so
for v in ${VARIANTS:-ct3 ct16 ct16m1}; do run $v 10 build/variants/libvlct_b200_$v.so; done
done
