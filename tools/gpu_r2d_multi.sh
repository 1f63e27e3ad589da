#!/bin/bash
# round-2 multi-GPU pass (8 GPUs of one box): H2D ceiling, weak / strong / ragged scaling points of bench.py
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
nvidia-smi topo -m > $O/r2d_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 $N; do
  [ $n -gt $N ] && continue
  timeout 120 $TR --nproc-per-node $n --master-port $((29500+n)) tools/h2d_ceiling.py 2>/dev/null | tail -1 >> $O/r2d_h2d.jsonl
done
cat $O/r2d_h2d.jsonl
B="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --sustain-seconds 2"
timeout 300 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N $B > $O/r2d_weak$N.json 2> $O/r2d_weak$N.err; echo "weak rc=$?"
timeout 300 $TR --nproc-per-node $N --master-port 29522 bench.py --gpus $N $B --scaling strong --batch 256 > $O/r2d_strong$N.json 2> $O/r2d_strong$N.err; echo "strong256 rc=$?"
timeout 300 $TR --nproc-per-node $N --master-port 29523 bench.py --gpus $N $B --scaling strong --batch 2048 > $O/r2d_strong2048_$N.json 2> $O/r2d_strong2048_$N.err; echo "strong2048 rc=$?"
timeout 300 $TR --nproc-per-node $N --master-port 29524 bench.py --gpus $N $B --ragged --vocab 44 > $O/r2d_ragged$N.json 2> $O/r2d_ragged$N.err; echo "ragged rc=$?"
timeout 300 $TR --nproc-per-node $N --master-port 29525 bench.py --gpus $N $B --ragged --vocab 44 --scaling strong --batch 2048 > $O/r2d_ragged_strong$N.json 2> $O/r2d_ragged_strong$N.err; echo "ragged strong rc=$?"
timeout 200 python bench.py --gpus 1 $B > $O/r2d_n1.json 2>$O/r2d_n1.err
timeout 200 python bench.py --gpus 1 $B --ragged --vocab 44 > $O/r2d_ragged1.json 2>$O/r2d_ragged1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "n", d["n_gpus"], d["scaling"], "value", d["value"], "e2e(i16)", d["e2e"]["value"], "e2e_f32", d["e2e_f32"]["value"], "sustained", (d.get("sustained") or {}).get("value"), "ms", d["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
