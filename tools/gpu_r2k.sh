#!/bin/bash
# round-2 GPU pass K: stride-2 depthwise with register prefetch; full GPU suite on the WRES GEMM
cd "$(dirname "$0")/.."
O=gpurun_out
for i in 1 2 3; do python tools/s2_time.py; V100_LIB=voice100_b200/libv100_prev.so python tools/s2_time.py; done > $O/r2k_s2.txt 2>&1; cat $O/r2k_s2.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2k_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2k_tests.log; tail -4 $O/r2k_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r2k_smoke.log
