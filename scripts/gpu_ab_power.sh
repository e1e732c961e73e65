#!/bin/bash
# A/B of library variants on one box: sustained ms + power of chosen kernel families
# usage: VARIANTS="a b" FAMILIES="edge,cell" bash scripts/gpu_ab_power.sh
mkdir -p gpurun_out
for V in default $VARIANTS; do
  if [ "$V" = default ]; then unset VLCT_B200_LIB; else export VLCT_B200_LIB=$PWD/build/variants/libvlct_b200_$V.so; fi
  echo "== $V"
  timeout 300 python scripts/gpu_power.py ${SIZE:-512} ${SECONDS_EACH:-2.5} ${FAMILIES:-edge} 2>&1 | grep family | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-8s %7.3f ms  %4.0f MHz %5.0f W' % (d['family'], d['ms_per_step'], d['sm_mhz'], d['power_instant_w']))"
done
