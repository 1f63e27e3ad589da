#!/bin/bash
# Build libv100 from another git revision into an alternate .so for same-box A/B runs:
#   tools/build_rev.sh <rev> voice100_b200/libv100_prev.so ; V100_LIB=voice100_b200/libv100_prev.so python bench.py ...
set -e
REV=$1; OUT=$(realpath -m $2); TMP=$(mktemp -d)
git archive $REV voice100_b200/csrc include | tar -x -C $TMP
cd $TMP/voice100_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --compiler-options -fPIC -shared -o $OUT $(ls *.cu)
rm -rf $TMP; echo built $OUT from $REV
