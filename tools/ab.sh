#!/bin/bash
# usage: tools/ab.sh libA.so libB.so   -> alternating bench runs on the same GPU
for i in 1 2; do for lib in "$@"; do
V100_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ra=d['roofline_all']
print('$lib', d['ms_per_step'], 'gemm', ra['gemm']['ms_per_step'], 'dw', ra['dwconv']['ms_per_step'], 'mel', ra['logmel']['ms_per_step'], 'clk', d['clocks']['sm_mhz'], [round(m,3) for k,m in d['launch_ms']][13:20])
"
done; done
