#!/bin/bash
# round-2 GPU pass J: weight half of every pair-GEMM stage by cp.async instead of TMA (ALDG)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv1x1 or convtranspose or conv_gemm" > $O/r2j_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2j_tests.log; tail -3 $O/r2j_tests.log
for i in 1 2; do for a in 1 0; do
  echo "== ALDG=$a"; V100_GEMM_ALDG=$a timeout 300 python tools/cublas_compare.py 2>&1 | grep pointwise
done; done > $O/r2j_gemm_ab.txt 2>&1; cat $O/r2j_gemm_ab.txt
for i in 1 2; do for a in 1 0; do
V100_GEMM_ALDG=$a timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2j_bench_a$a.$i.json 2>$O/r2j_bench_a$a.$i.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2j_bench_a$a.$i.json").read().strip().splitlines()[-1])
    print("ALDG=$a run $i ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k:v["ms_per_step"] for k,v in d["roofline_all"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("bench ERR", e, open("gpurun_out/r2j_bench_a$a.$i.err").read()[-800:])
PY
done; done
PROF_WHICH=gemm timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_pair" -f -o $O/r2j_prof_gemm python tools/profile_kernels.py > $O/r2j_ncu.log 2>&1; echo "ncu rc=$?"
