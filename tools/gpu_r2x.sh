#!/bin/bash
# round-2 GPU pass X (final): full validation and measurements on the WRES GEMM + bulk-staged depthwise code
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rP > $O/r2x_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2x_tests.log; tail -3 $O/r2x_tests.log; grep -c PASSED $O/r2x_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2x_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2x_smoke.log; tail -3 $O/r2x_smoke.log
timeout 300 python tools/cublas_compare.py > $O/r2x_cublas.txt 2>&1; cat $O/r2x_cublas.txt
(timeout 300 python tools/dw_time.py; timeout 100 python tools/s2_time.py; timeout 300 python tools/logmel_time.py 2>&1 | tail -2) > $O/r2x_kernel_times.txt 2>&1; cat $O/r2x_kernel_times.txt
for i in 1 2; do
timeout 500 python bench.py --steps 20 --warmup 5 > $O/r2x_bench_full$i.json 2> $O/r2x_bench_full$i.err; echo "full bench $i rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2x_bench_full$i.json").read().strip().splitlines()[-1])
    print("run $i ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "e2e_f32", d["e2e_f32"]["value"], "sustained", d["sustained"]["value"], d["step_model"], {k:v["ms_per_step"] for k,v in d["roofline_all"].items()}, d["roofline"]["frac"], d["clocks"])
except Exception as e: print("ERR", e, open("gpurun_out/r2x_bench_full$i.err").read()[-600:])
PY
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2x_bench_ref.json 2> $O/r2x_bench_ref.err; tail -c 400 $O/r2x_bench_ref.json
timeout 300 python bench.py --steps 20 --warmup 5 --dtype f16 --no-cpu-baseline --no-gpu-eager > $O/r2x_bench_f16.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench_f16.json').read().strip().splitlines()[-1]); print('f16 ms/step', d['ms_per_step'], 'value', d['value'], 'sustained', d['sustained']['value'])"
timeout 300 python bench.py --steps 20 --warmup 5 --ragged --vocab 44 --no-cpu-baseline --no-gpu-eager > $O/r2x_bench_ragged.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench_ragged.json').read().strip().splitlines()[-1]); print('ragged ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])"
for wl in tts asr_v2 tts_v2; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 > $O/r2x_bench_$wl.json 2>$O/r2x_bench_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['ms_per_step'], 'ms', d['value'], d['unit'], 'e2e', d['e2e']['value'], 'launches/step', d.get('gpu_launches_per_step'))" || tail -3 $O/r2x_bench_$wl.err; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"logmel|conv_gemm|dw_|ctc_finalize|expand_dw" -c 150 --csv --log-file $O/r2x_step_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2x_ncu_bench.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dw_mma|dw_bulk|dw_s2|conv_gemm|logmel" -f -o $O/r2x_prof_full python tools/profile_kernels.py > $O/r2x_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O | grep r2m | head -40
