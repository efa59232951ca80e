#!/bin/bash
# final build at 8 GPUs: one cfg2 line (weak scaling), dp_parity included
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus 8 --steps 10 --warmup 3 --lite > gpurun_out/final_n8_cfg2.json 2> gpurun_out/final_n8_cfg2.err
echo "bench N=8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/final_n8_cfg2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','phase_ms_sequential']}, (d.get('dp_parity') or {}).get('ok'))
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/final_n8_cfg2.err').read()[-2500:])
PY
