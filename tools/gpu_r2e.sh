#!/bin/bash
# round-2 GPU pass E: dw staging fix, library comparison, sanitizers, step launch list + DRAM bytes, full captures
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -rP > $O/r2e_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2e_tests.log; tail -3 $O/r2e_tests.log
timeout 300 python tools/dw_time.py > $O/r2e_dw_time.txt 2>&1; cat $O/r2e_dw_time.txt
timeout 300 python tools/logmel_time.py 2>&1 | tail -1
timeout 300 python tools/cublas_compare.py > $O/r2e_cublas.txt 2>&1; cat $O/r2e_cublas.txt
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2e_bench$i.json 2>$O/r2e_bench$i.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_bench$i.json").read().strip().splitlines()[-1])
    print("bench$i ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["step_model"], {k:v["ms_per_step"] for k,v in d["roofline_all"].items()})
except Exception as e: print("bench$i ERR", e, open("gpurun_out/r2e_bench$i.err").read()[-800:])
PY
done
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/r2e_san_memcheck.log 2>&1; tail -3 $O/r2e_san_memcheck.log
SAN_MODE=nogemm timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/r2e_san_racecheck.log 2>&1; tail -2 $O/r2e_san_racecheck.log
SAN_MODE=nogemm timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_small.py > $O/r2e_san_synccheck.log 2>&1; tail -2 $O/r2e_san_synccheck.log
SAN_MODE=nogemm timeout 600 compute-sanitizer --tool initcheck python tools/sanitize_small.py > $O/r2e_san_initcheck.log 2>&1; tail -2 $O/r2e_san_initcheck.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"logmel|conv_gemm|dw_|ctc_finalize|expand_dw" -c 150 --csv --log-file $O/r2e_step_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2e_ncu_bench.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dw_mma|dw_s2|conv_gemm|logmel" -f -o $O/r2e_prof_full python tools/profile_kernels.py > $O/r2e_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O | grep r2e
