#!/bin/bash
# round-2 GPU pass S: 128-column pair GEMM for short rows, half-tile skip in the bulk depthwise kernel (TTS shapes)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2s_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2s_tests.log; tail -4 $O/r2s_tests.log
timeout 300 python tools/tts_prof.py > $O/r2s_tts_prof.txt 2>&1; tail -42 $O/r2s_tts_prof.txt
for wl in tts tts_v2 asr_v2; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 > $O/r2s_bench_$wl.json 2>$O/r2s_bench_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['ms_per_step'], 'ms', d['value'], d['unit'], 'e2e', d['e2e']['value'])" || tail -3 $O/r2s_bench_$wl.err; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('asr ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])"
