#!/bin/bash
# round-2 GPU pass H: weights-resident (TMEM) pair GEMM + compile-time epilogue variants
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv1x1 or convtranspose or conv_gemm" > $O/r2h_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2h_tests.log; tail -5 $O/r2h_tests.log
for i in 1 2; do
  echo "== WRES=1"; timeout 300 python tools/cublas_compare.py 2>&1 | grep pointwise
  echo "== WRES=0"; V100_GEMM_WRES=0 timeout 300 python tools/cublas_compare.py 2>&1 | grep pointwise
  echo "== prev (HEAD)"; V100_LIB=voice100_b200/libv100_prev.so timeout 300 python tools/cublas_compare.py 2>&1 | grep pointwise
done > $O/r2h_gemm_ab.txt 2>&1; cat $O/r2h_gemm_ab.txt
for i in 1 2; do for v in 1 0; do
V100_GEMM_WRES=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2h_bench_w$v.$i.json 2>$O/r2h_bench_w$v.$i.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h_bench_w$v.$i.json").read().strip().splitlines()[-1])
    print("WRES=$v run $i ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k:v["ms_per_step"] for k,v in d["roofline_all"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("bench ERR", e, open("gpurun_out/r2h_bench_w$v.$i.err").read()[-800:])
PY
done; done
PROF_WHICH=gemm timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm" -f -o $O/r2h_prof_gemm python tools/profile_kernels.py > $O/r2h_ncu.log 2>&1; echo "ncu rc=$?"
