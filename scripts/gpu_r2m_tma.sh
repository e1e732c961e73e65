#!/bin/bash
# round 2: k_edge_efield_tma (inputs staged by cp.async.bulk) -- parity first
# (short timeouts: a lost mbarrier phase would hang), then same-box A/B
mkdir -p gpurun_out
TAG=${TAG:-r2m}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 > gpurun_out/pytest_parity_$TAG.log 2>&1; rc=$?; echo "parity rc=$rc"; tail -5 gpurun_out/pytest_parity_$TAG.log
if [ $rc != 0 ]; then exit 0; fi
if [ "${TESTS:-1}" = 1 ]; then
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -8 gpurun_out/pytest_gpu_$TAG.log
fi
RUNS=${RUNS:-"base:6:enzo-e_b200/csrc/libvlct_b200.so tma:14:enzo-e_b200/csrc/libvlct_b200.so"} TAG=$TAG bash scripts/gpu_ab_mask.sh
