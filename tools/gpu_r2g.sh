#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for i in 1 2; do
python tools/dw_time.py; V100_LIB=voice100_b200/libv100_pad16.so python tools/dw_time.py
done > $O/r2g_dw_ab.txt 2>&1; cat $O/r2g_dw_ab.txt
V100_LIB=voice100_b200/libv100_pad16.so timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "dwconv" 2>&1 | tail -2
PROF_WHICH=dw timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dw_mma_kernel<6" -c 1 -f -o $O/r2g_prof_q6 python tools/profile_kernels.py > $O/r2g_ncu.log 2>&1; echo "ncu rc=$?"
