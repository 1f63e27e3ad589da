#!/bin/bash
# round-2 GPU pass I: L2 prefetch distance of the GEMM producer
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv1x1 or convtranspose or conv_gemm" > $O/r2i_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2i_tests.log; tail -3 $O/r2i_tests.log
for pf in 0 4 8 16 32 0 8; do
  echo "== PF=$pf"; V100_GEMM_PF=$pf timeout 300 python tools/cublas_compare.py 2>&1 | grep pointwise | cut -d'|' -f1
done > $O/r2i_gemm_pf.txt 2>&1; cat $O/r2i_gemm_pf.txt
for i in 1 2; do for pf in 8 0; do
V100_GEMM_PF=$pf timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2i_bench_pf$pf.$i.json 2>$O/r2i_bench_pf$pf.$i.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2i_bench_pf$pf.$i.json").read().strip().splitlines()[-1])
    print("PF=$pf run $i ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k:v["ms_per_step"] for k,v in d["roofline_all"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("bench ERR", e, open("gpurun_out/r2i_bench_pf$pf.$i.err").read()[-800:])
PY
done; done
