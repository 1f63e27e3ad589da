#!/bin/bash
# round-2 GPU pass F: stride-2 polyphase kernel, PDL A/B, step launch list with DRAM bytes, secondary workloads
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -rP > $O/r2f_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2f_tests.log; tail -3 $O/r2f_tests.log; grep -c PASSED $O/r2f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2f_smoke.log; tail -3 $O/r2f_smoke.log
python - <<'PY'
import sys, torch, math
sys.path.insert(0, "/root/repo")
from voice100_b200 import kernels as K
dev="cuda"
x = K.empty_ncw(256, 256, 1501, dev); x.data.normal_()
w = (torch.randn(256, 11, device=dev)/3).to(torch.bfloat16)
s, b = torch.ones(256, device=dev), torch.zeros(256, device=dev)
for simt in (False, True):
    for _ in range(3): K.dwconv(x, w, s, b, 11, 2, 1, simt=simt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): K.dwconv(x, w, s, b, 11, 2, 1, simt=simt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    gb = 256*256*(1501+751)*2/1e9
    print(f"dw stride 2 (256 ch, 1501 -> 751, k=11) {'simt' if simt else 'polyphase mma'}: {ms*1e3:.1f} us = {gb/ms:.0f} GB/s")
PY
for i in 1 2 3; do for pdl in 1 0; do V100_PDL=$pdl timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2f_bench_pdl$pdl.$i.json 2>$O/r2f_bench.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2f_bench_pdl$pdl.$i.json").read().strip().splitlines()[-1])
    print("PDL=$pdl run $i ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["step_model"], {k:v["ms_per_step"] for k,v in d["roofline_all"].items()})
except Exception as e: print("ERR", e, open("gpurun_out/r2f_bench.err").read()[-600:])
PY
done; done
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2f_bench_full.json 2> $O/r2f_bench_full.err; echo "full bench rc=$?"; tail -c 600 $O/r2f_bench_full.json
timeout 300 python bench.py --steps 20 --warmup 5 --dtype f16 --no-cpu-baseline --no-gpu-eager > $O/r2f_bench_f16.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench_f16.json').read().strip().splitlines()[-1]); print('f16 ms/step', d['ms_per_step'], 'value', d['value'], 'sustained', d['sustained']['value'])"
for wl in tts asr_v2 tts_v2; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 > $O/r2f_bench_$wl.json 2>$O/r2f_bench_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['ms_per_step'], 'ms', d['value'], d['unit'], 'e2e', d['e2e']['value'], 'launches/step', d.get('gpu_launches_per_step'))" || tail -3 $O/r2f_bench_$wl.err; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"logmel|conv_gemm|dw_|ctc_finalize|expand_dw" -c 150 --csv --log-file $O/r2f_step_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --sustain-seconds 0 > $O/r2f_ncu_bench.log 2>&1; echo "ncu step rc=$?"
PROF_WHICH=dw timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dw_s2" -c 1 -f -o $O/r2f_prof_s2 python tools/profile_kernels.py > $O/r2f_ncu_s2.log 2>&1; echo "ncu s2 rc=$?"
ls -la $O | grep r2f | head -30
