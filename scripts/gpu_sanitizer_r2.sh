#!/bin/bash
# compute-sanitizer over the round-2 kernels (pair kernels, k_edge_efield_tma,
# k_ct_tma forced on a small block with 3 x 3 tiles and four z chunks)
mkdir -p gpurun_out
OUT=gpurun_out/r2_sanitizer.txt
: > $OUT
SEL='pair_kernels and tiles_and_chunks and default and (mhd_hlld_plm] or mhd_hlld_athena_de)'
for TOOL in memcheck synccheck; do
  echo "== compute-sanitizer --tool $TOOL python -m pytest tests/test_gpu_parity.py -m gpu -k \"$SEL\"" >> $OUT
  timeout 420 compute-sanitizer --tool $TOOL --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 400 -k "$SEL" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|Barrier error|=========     at|error" | head -30 >> $OUT
  echo "rc=$?" >> $OUT
done
cat $OUT
