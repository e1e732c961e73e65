#!/bin/bash
# ncu --set full (+ source) of every kernel of one folded 512^3 step, and of the
# scalar kernels of the 4-scalar workload. The reports are summarised ON the
# GPU box (gpurun brings back at most 64 MiB): raw-page summary per launch and
# the dynamic opcode mix / stall attribution of the PLM z march, the edge kernel
# and the CFL-folding update; only the small scalar report travels whole.
mkdir -p gpurun_out /tmp/ncu
TAG=${TAG:-r2h}
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:k_flux|k_edge|k_update|k_face|k_timestep_shell' -s 39 -c 13 \
  -f -o /tmp/ncu/prof_step512_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_step512_$TAG.log 2>&1; echo "ncu step rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scalar_update -s 6 -c 2 \
  -f -o gpurun_out/prof_scalar512_$TAG python bench.py --workload ot_s4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_scalar512_$TAG.log 2>&1; echo "ncu scalar rc=$?"
python scripts/ncu_summary.py /tmp/ncu/prof_step512_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_full_step_512.txt 2>&1
python scripts/ncu_summary.py gpurun_out/prof_scalar512_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_full_scalar_512.txt 2>&1
for K in "k_flux_march<2, 1" "k_edge_efield" "k_update<1, 0, 0, 1" "k_flux_x<1"; do
  N=$(echo "$K" | tr -c 'a-z0-9_' '_')
  ncu -i /tmp/ncu/prof_step512_$TAG.ncu-rep --page source --csv --kernel-name "regex:$(echo $K | cut -d'<' -f1)" --launch-count 1 ${SKIP:-} > /tmp/ncu/src_$N.csv 2>/dev/null
  python scripts/ncu_opmix.py /tmp/ncu/src_$N.csv > gpurun_out/${TAG}_opmix_$N.txt 2>&1
done
ls -la /tmp/ncu gpurun_out | tail -20
