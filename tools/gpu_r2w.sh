#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "dwconv or tts or v2 or layout" > $O/r2w_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2w_tests.log; tail -6 $O/r2w_tests.log
echo "== rows kernel"; timeout 300 python tools/tts_prof.py 2>&1 | grep -E "dwconv|sum"
echo "== bulk kernel"; V100_DW_ROWS=0 timeout 300 python tools/tts_prof.py 2>&1 | grep -E "dwconv|sum"
timeout 300 python bench.py --workload tts --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tts', d['ms_per_step'], 'ms', d['value'], d['config']['launch'], d['config']['eager_ms_per_step'])"
