#!/bin/bash
# round 2: k_ct_tma (edge E + face B, TMA-staged) -- parity with the kernel
# forced on (VLCT_PAIR_MASK=30), short timeouts, then same-box A/B
mkdir -p gpurun_out
TAG=${TAG:-r2q}
VLCT_PAIR_MASK=30 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parts.py tests/test_gpu_fused_timestep.py -m gpu -q -x --timeout 100 > gpurun_out/pytest_parity_$TAG.log 2>&1; rc=$?; echo "parity rc=$rc"; tail -15 gpurun_out/pytest_parity_$TAG.log
if [ $rc != 0 ]; then exit 0; fi
if [ "${TESTS:-1}" = 1 ]; then
VLCT_PAIR_MASK=30 timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -8 gpurun_out/pytest_gpu_$TAG.log
fi
RUNS=${RUNS:-"base:14:enzo-e_b200/csrc/libvlct_b200.so ct:30:enzo-e_b200/csrc/libvlct_b200.so"} TAG=$TAG bash scripts/gpu_ab_mask.sh
