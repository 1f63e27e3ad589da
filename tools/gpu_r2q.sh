#!/bin/bash
# round-2 GPU pass Q: new dw_bulk edge/determinism test, sanitizers over the final kernels
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "dwconv" > $O/r2q_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2q_tests.log; tail -5 $O/r2q_tests.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/r2q_san_memcheck.log 2>&1; tail -3 $O/r2q_san_memcheck.log
SAN_MODE=nogemm timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/r2q_san_racecheck.log 2>&1; tail -3 $O/r2q_san_racecheck.log
SAN_MODE=nogemm timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_small.py > $O/r2q_san_synccheck.log 2>&1; tail -3 $O/r2q_san_synccheck.log
SAN_MODE=nogemm timeout 600 compute-sanitizer --tool initcheck python tools/sanitize_small.py > $O/r2q_san_initcheck.log 2>&1; tail -3 $O/r2q_san_initcheck.log
SAN_MODE=v2 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/r2q_san_memcheck_v2.log 2>&1; tail -3 $O/r2q_san_memcheck_v2.log
