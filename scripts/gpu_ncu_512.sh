#!/bin/bash
# ncu --set full of one launch of each flux kernel at 512^3 (kernel order in a
# step: x_nn y_nn z_nn edge face update x_plm y_plm z_plm ...)
mkdir -p gpurun_out
TAG=${TAG:-r1d}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_flux -s 24 -c 6 \
  -f -o gpurun_out/prof_flux512_$TAG python bench.py --size 512 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_flux512_$TAG.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*512*.ncu-rep
