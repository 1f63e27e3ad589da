#!/bin/bash
# round-2 GPU pass B: fused expand+depthwise kernel
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "expand_dw" > $O/r2b_fused_tests.log 2>&1; echo "fused pytest rc=$?"; tail -15 $O/r2b_fused_tests.log
timeout 300 python tools/fused_time.py > $O/r2b_fused_time.txt 2>&1; cat $O/r2b_fused_time.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -rP > $O/r2b_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2b_tests.log; tail -8 $O/r2b_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2b_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2b_smoke.log; tail -3 $O/r2b_smoke.log
timeout 300 python tools/logmel_time.py 2>&1 | tail -1
for f in 1 0 1 0; do V100_FUSE=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2b_bench_fuse$f.json 2>$O/r2b_bench_fuse$f.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b_bench_fuse$f.json").read().strip().splitlines()[-1])
    print("FUSE=$f ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["step_model"], {k:v["ms_per_step"] for k,v in d["roofline_all"].items()})
except Exception as e: print("FUSE=$f ERR", e, open("gpurun_out/r2b_bench_fuse$f.err").read()[-800:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_dw -c 2 -f -o $O/r2b_prof_fused python tools/fused_time.py --once > $O/r2b_ncu_fused.log 2>&1; echo "ncu fused rc=$?"
ls -la $O | grep r2b
