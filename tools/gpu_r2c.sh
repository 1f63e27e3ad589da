#!/bin/bash
# round-2 GPU pass C: new depthwise kernel (shared data fragments), full tests, bench A/B, profiles
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "dwconv" > $O/r2c_dw_tests.log 2>&1; echo "dw pytest rc=$?"; tail -4 $O/r2c_dw_tests.log
timeout 300 python tools/dw_time.py > $O/r2c_dw_time.txt 2>&1; cat $O/r2c_dw_time.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -rP > $O/r2c_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2c_tests.log; tail -4 $O/r2c_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c_smoke.log; tail -3 $O/r2c_smoke.log
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2c_bench$i.json 2>$O/r2c_bench$i.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c_bench$i.json").read().strip().splitlines()[-1])
    print("bench$i ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["step_model"], {k:v["ms_per_step"] for k,v in d["roofline_all"].items()})
    print([m for k,m in d["launch_ms"]])
except Exception as e: print("bench$i ERR", e, open("gpurun_out/r2c_bench$i.err").read()[-800:])
PY
done
PROF_WHICH=dw timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_mma -c 6 -f -o $O/r2c_prof_dw python tools/profile_kernels.py > $O/r2c_ncu_dw.log 2>&1; echo "ncu dw rc=$?"
ls -la $O | grep r2c
