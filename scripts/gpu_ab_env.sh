#!/bin/bash
# same-box A/B over environment settings: RUNS="name:VAR=value,VAR2=value ..."
mkdir -p gpurun_out
TAG=${TAG:-abe}
for rep in 1 2; do
for r in $RUNS; do
  name=${r%%:*}; envs=${r#*:}
  env $(echo $envs | tr ',' ' ') timeout 600 python bench.py --workload ${WL:-ot} --steps 6 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$name.json").read().strip().splitlines()[-1])
    print("$name ms/step %.2f " % d["ms_per_step"], d["clocks"].get("sm_mhz"), {k[2:]: round(v["ms_per_step"], 2) for k,v in d["kernels"].items() if v["ms_per_step"] > 0.1})
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/bench_${TAG}_$name.err").read()[-1500:])
PY
done
done
