#!/bin/bash
# Multi-GPU visit (`gpurun --gpus 8 -- 'bash tools/gpu_scale.sh'`, or `--gpus 2 -- 'bash tools/gpu_scale.sh 2'`): the driver's
# SCALE line at the largest N (full bench), the smaller N with --lite, and cfg5 (BASELINE configs[4], strong scaling).
mkdir -p gpurun_out
NS=${@:-8 4}
run() {  # name, nproc, extra args
  if [ "$2" = "1" ]; then timeout 600 python bench.py --steps 10 --warmup 3 $3 > gpurun_out/$1.json 2> gpurun_out/$1.err
  else timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus $2 --steps 10 --warmup 3 $3 > gpurun_out/$1.json 2> gpurun_out/$1.err; fi
  echo "$1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$1.json').read().strip().splitlines()[-1])
    print('$1', {k:d[k] for k in ['value','ms_per_step','phase_ms_sequential']}, 'dp_ok', (d.get('dp_parity') or {}).get('ok'), 'e2e', (d.get('e2e') or {}).get('value'))
except Exception as e:
    print('no bench line $1', e); print(open('gpurun_out/$1.err').read()[-2500:])
PY
}
first=1
for N in $NS; do
  if [ $first = 1 ]; then run scale_n$N $N ""; first=0; else run scale_n$N $N "--lite"; fi
  run cfg5_n$N $N "--lite --workload cfg5"
done
