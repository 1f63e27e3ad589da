#!/bin/bash
# epilogue clean-up A/B (incremental tile coordinates, shared-space stores)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv1x1 or convtranspose or conv_gemm" > $O/r2n_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2n_tests.log; tail -3 $O/r2n_tests.log
for i in 1 2; do
  echo "== new"; timeout 300 python tools/cublas_compare.py 2>&1 | grep pointwise | cut -d'|' -f1
  echo "== prev"; V100_LIB=voice100_b200/libv100_prev.so timeout 300 python tools/cublas_compare.py 2>&1 | grep pointwise | cut -d'|' -f1
done > $O/r2n_gemm_ab.txt 2>&1; cat $O/r2n_gemm_ab.txt
