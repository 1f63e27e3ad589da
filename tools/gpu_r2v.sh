#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "dwconv or tts or asr_matches or v2" > $O/r2v_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2v_tests.log; tail -4 $O/r2v_tests.log
timeout 300 python tools/tts_prof.py 2>&1 | grep -E "dwconv|sum"
timeout 300 python tools/dw_time.py
for wl in tts; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', d['ms_per_step'], 'ms', d['value'], d['unit'], 'e2e', d['e2e']['value'])"; done
