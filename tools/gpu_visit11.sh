#!/bin/bash
# 2-GPU visit: N=2 bench after the vectorised in-kernel exchange; cfg5 (4096 x 256, strong scaling) at N = 1, 2
mkdir -p gpurun_out
run() {  # name, nproc, extra args
  if [ "$2" = "1" ]; then timeout 600 python bench.py --steps 10 --warmup 3 --lite $3 > gpurun_out/$1.json 2> gpurun_out/$1.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus $2 --steps 10 --warmup 3 --lite $3 > gpurun_out/$1.json 2> gpurun_out/$1.err; fi
  echo "$1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$1.json').read().strip().splitlines()[-1])
    print('$1', {k:d[k] for k in ['value','ms_per_step','phase_ms_sequential']}, (d.get('dp_parity') or {}).get('ok'))
except Exception as e:
    print('no bench line $1', e); print(open('gpurun_out/$1.err').read()[-2500:])
PY
}
run bench_n2 2 ""
run cfg5_n1 1 "--workload cfg5"
run cfg5_n2 2 "--workload cfg5"
