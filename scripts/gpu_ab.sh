#!/bin/bash
# parity tests, then A/B of library variants given in $VARIANTS (paths), bench at $SIZES
mkdir -p gpurun_out
TAG=${TAG:-ab}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for lib in ${VARIANTS:-enzo-e_b200/csrc/libvlct_b200.so}; do
  for size in ${SIZES:-256}; do
    name=$(basename $lib .so)
    VLCT_B200_LIB=$PWD/$lib timeout 900 python bench.py --size $size --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_${name}_$size.json 2> gpurun_out/bench_${TAG}_${name}_$size.err
    echo "$name $size rc=$?"
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_${name}_$size.json"))
    print("  value %.4g ms/step %.2f compute_only %.2f" % (d["value"], d["ms_per_step"], d["compute_only_ms"]))
    for k,v in d["kernels"].items(): print("   %-22s %8.3f ms x%g" % (k, v["ms_per_step"], v["launches_per_step"]))
except Exception as e: print("  failed", e)
PY
  done
done
