#!/bin/bash
# Multi-GPU scaling visit (run with `gpurun --gpus 8`): the full bench line at every N the driver's SCALE run uses
mkdir -p gpurun_out
for N in ${@:-4 8}; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29720+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  echo "bench N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_n$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','phase_ms_sequential','dp_parity']})
    print('e2e', d['e2e']['value'])
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/scale_n$N.err').read()[-2500:])
PY
done
