#!/bin/bash
# 8-GPU visit: the driver's SCALE line at N=8 (full bench), N=4 (lite), cfg5 strong scaling at N=4, 8 (lite)
mkdir -p gpurun_out
run() {  # name, nproc, extra args
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus $2 --steps 10 --warmup 3 $3 > gpurun_out/$1.json 2> gpurun_out/$1.err
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
run scale_n8 8 ""
run cfg5_n8 8 "--lite --workload cfg5"
run scale_n4 4 "--lite"
run cfg5_n4 4 "--lite --workload cfg5"
