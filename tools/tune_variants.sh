#!/bin/bash
# Build tuning variants of libvicres.so (CTA size, blocks/SM, phase barriers) into _variants/.
# usage: tools/tune_variants.sh name "flags" ...
set -e
cd "$(dirname "$0")/.."
BASE="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -ffp-contract=off -fmad=false -DVR_PHASE_BARRIERS"
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc $BASE $flags -o _variants/lib_$name.so vicres_b200/csrc/vicres_capi.cu vicres_b200/csrc/vicres_route.cu vicres_b200/csrc/vicres_atmos.cu vicres_b200/csrc/vicres_files.cpp &
done
wait
ls -la _variants/
