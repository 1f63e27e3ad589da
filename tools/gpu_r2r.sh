#!/bin/bash
cd "$(dirname "$0")/.."
for pr in 74 44 37 30; do V100_GEMM_PAIRS=$pr timeout 200 python tools/overlap_probe.py; done 2>&1 | tee gpurun_out/r2r_overlap.txt
