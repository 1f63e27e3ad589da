#!/bin/bash
# round-2 GPU pass L: depthwise with bulk-copy staging (dw_bulk_kernel) vs dw_mma_kernel
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "dwconv" > $O/r2l_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2l_tests.log; tail -5 $O/r2l_tests.log
for i in 1 2; do
python tools/dw_time.py | sed 's/^default/bulk/'; V100_DW_BULK=0 python tools/dw_time.py | sed 's/^default/split/'
done > $O/r2l_dw_ab.txt 2>&1; cat $O/r2l_dw_ab.txt
PROF_WHICH=dw timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dw_bulk_kernel" -c 4 -f -o $O/r2l_prof_dw python tools/profile_kernels.py > $O/r2l_ncu.log 2>&1; echo "ncu rc=$?"
