#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for i in 1 2 3 4; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2t_bench$i.json 2>$O/r2t_bench$i.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2t_bench$i.json").read().strip().splitlines()[-1])
    print("run $i ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "ramp", d["config"]["clock_ramp_steps"], d["clocks"])
except Exception as e: print("ERR", e, open("gpurun_out/r2t_bench$i.err").read()[-600:])
PY
sleep 5
done
