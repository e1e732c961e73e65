#!/bin/bash
# A/B of library variants given in $VARIANTS (paths) without the test suite, bench at $SIZES
mkdir -p gpurun_out
TAG=${TAG:-ab}
for rep in 1 2; do
for lib in ${VARIANTS:-enzo-e_b200/csrc/libvlct_b200.so}; do
  for size in ${SIZES:-512}; do
    name=$(basename $lib .so)
    VLCT_B200_LIB=$PWD/$lib timeout 900 python bench.py --size $size --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_${name}_$size.json 2> gpurun_out/bench_${TAG}_${name}_$size.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_${name}_$size.json"))
    print("$name $size ms/step %.2f " % d["ms_per_step"], {k[2:]: round(v["ms_per_step"], 2) for k,v in d["kernels"].items() if v["ms_per_step"] > 0.1})
except Exception as e: print("$name failed", e)
PY
  done
done
done
