#!/bin/bash
# round 2, last evidence run: the new parity shape + smoke, ncu --set full of
# every kernel of one final 512^3 step (summarised on the box), board power per
# kernel family with the final kernels
mkdir -p gpurun_out /tmp/ncu
TAG=${TAG:-r2f2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "pair_kernels" > gpurun_out/pytest_pair_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_pair_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; grep "smoke ok" gpurun_out/smoke_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_flux|k_ct_tma|k_update|k_timestep_boxes' -s 33 -c 11 \
  -f -o /tmp/ncu/prof_step512_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_step512_$TAG.log 2>&1; echo "ncu step rc=$?"
python scripts/ncu_summary.py /tmp/ncu/prof_step512_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_full_step_512.txt 2>&1
for K in "k_ct_tma" "k_update2"; do
  ncu -i /tmp/ncu/prof_step512_$TAG.ncu-rep --page source --csv --kernel-name "regex:$K" --launch-count 1 > /tmp/ncu/src_$K.csv 2>/dev/null
  python scripts/ncu_opmix.py /tmp/ncu/src_$K.csv > gpurun_out/${TAG}_opmix_$K.txt 2>&1
done
timeout 300 python scripts/gpu_power.py 512 3.0 all,flux,ct,update 2>&1 | grep family > gpurun_out/${TAG}_power.jsonl; cat gpurun_out/${TAG}_power.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-8s %7.3f ms  %4.0f MHz %5.0f W' % (d['family'], d['ms_per_step'], d['sm_mhz'], d['power_instant_w']))"
grep -c "gpu__time_duration" gpurun_out/${TAG}_ncu_full_step_512.txt
