#!/bin/bash
# build an A/B variant of the CUDA library: build_variant.sh NAME "-DFLAG ..."
# -> build/variants/libvlct_b200_NAME.so (select with VLCT_B200_LIB=...)
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd $(dirname $0)/.. && pwd)
OUT=$ROOT/build/variants; mkdir -p $OUT/obj_$NAME
cd $ROOT/enzo-e_b200/csrc
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr $FLAGS"
for f in ${FILES:-vlct_kernels.cu}; do $NV -c $f -o $OUT/obj_$NAME/${f%.cu}.o & done
wait
OBJS=""
for f in vlct_flux vlct_kernels vlct_api vlct_config vlct_selftest; do
  if [ -f $OUT/obj_$NAME/$f.o ]; then OBJS="$OBJS $OUT/obj_$NAME/$f.o"; else OBJS="$OBJS $f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libvlct_b200_$NAME.so $OBJS -lcudart
echo built $OUT/libvlct_b200_$NAME.so
