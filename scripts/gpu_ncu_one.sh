#!/bin/bash
# ncu --set full of ONE launch of a kernel chosen by regex ($KERNEL), skip $SKIP matches
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_flux_march} -s ${SKIP:-6} -c 1 \
  -f -o gpurun_out/prof_one_${TAG:-x} python bench.py --size ${SIZE:-256} --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_one_${TAG:-x}.log 2>&1; echo "ncu rc=$?"
