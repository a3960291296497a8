#!/bin/bash
# Experiment builds of libia_b200 with other compiler flags (NOT the product build: __graft_entry__.build()).
#   scripts/build_variant.sh fmad   ->  instantavatar_b200/libia_b200_fmad.so   (-fmad=true: free FMA contraction; results are
#                                        then no longer bit-identical to the CPU oracle -- used to measure what the explicit-FMA
#                                        discipline costs, see scripts/mlp_share.py / profiles/)
set -e
cd "$(dirname "$0")/.."
case "$1" in
  fmad) FLAGS="-fmad=true"; OUT=instantavatar_b200/libia_b200_fmad.so ;;
  *) echo "usage: $0 fmad"; exit 2 ;;
esac
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a $FLAGS -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
  -o $OUT instantavatar_b200/csrc/*.cu
echo built $OUT
