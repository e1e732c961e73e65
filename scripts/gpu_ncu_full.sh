#!/bin/bash
# ncu --set full of the flux / edge / update kernels (256^3 so replays stay short)
mkdir -p gpurun_out
SIZE=${SIZE:-256}
TAG=${TAG:-r1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flux -s 9 -c 3 \
  -f -o gpurun_out/prof_flux_$TAG python bench.py --size $SIZE --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_flux_$TAG.log 2>&1; echo "ncu flux rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_update|k_edge|k_face|k_timestep' -s 8 -c 4 \
  -f -o gpurun_out/prof_rest_$TAG python bench.py --size $SIZE --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rest_$TAG.log 2>&1; echo "ncu rest rc=$?"
ls -la gpurun_out/*.ncu-rep
