#!/bin/bash
cd "$(dirname "$0")/.."
for pr in 74 56 37 18; do echo "== pairs $pr"; V100_GEMM_PAIRS=$pr timeout 300 python tools/cublas_compare.py 2>&1 | grep "2048->512\|1024->256\|1024->512" | cut -d'|' -f1; done
