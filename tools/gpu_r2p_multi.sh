#!/bin/bash
# round-2 final multi-GPU point: weak scaling at N GPUs with the final kernels (int16 and fp32 e2e), plus N=1 on the same box
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 2"
timeout 200 python bench.py --gpus 1 $B > $O/r2p_n1.json 2>$O/r2p_n1.err
timeout 300 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N $B > $O/r2p_weak$N.json 2> $O/r2p_weak$N.err; echo "weak rc=$?"
timeout 300 $TR --nproc-per-node $N --master-port 29522 bench.py --gpus $N $B --scaling strong --batch 256 > $O/r2p_strong$N.json 2> $O/r2p_strong$N.err; echo "strong256 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2p_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "n", d["n_gpus"], d["scaling"], "value", d["value"], "e2e(i16)", d["e2e"]["value"], "e2e_f32", d["e2e_f32"]["value"], "sustained", (d.get("sustained") or {}).get("value"), "ms", d["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
