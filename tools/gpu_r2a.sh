#!/bin/bash
# round-2 GPU pass A: tests, smoke, log-mel A/B, bench with/without PDL, launch list, log-mel ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -rP -x > $O/r2a_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2a_tests.log
tail -5 $O/r2a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2a_smoke.log; tail -4 $O/r2a_smoke.log
for nw in 16 8; do V100_MEL_WARPS=$nw timeout 120 python tools/logmel_time.py 2>&1 | tail -1 | sed "s/^/NW=$nw /"; done > $O/r2a_logmel.txt 2>&1; cat $O/r2a_logmel.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"; tail -c 1500 $O/r2a_bench.json
V100_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2a_bench_nopdl.json 2> $O/r2a_bench_nopdl.err; echo "bench nopdl rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2a_bench_pdl2.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2a_bench.json","r2a_bench_nopdl.json","r2a_bench_pdl2.json"):
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "e2e_f32", d["e2e_f32"]["value"], "sum", d["step_model"], {k:v["ms_per_step"] for k,v in d["roofline_all"].items()})
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2a_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel -s 2 -c 1 -f -o $O/r2a_prof_mel python tools/logmel_time.py > $O/r2a_ncu_mel.log 2>&1; echo "ncu mel rc=$?"
ls -la $O | tail -15
