#!/bin/bash
# 2-GPU check of the final build: DP parity tests + N=2 bench (cfg2 and cfg5, lite)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1
echo "dp rc=$?"; tail -3 gpurun_out/pytest_dp.log
for wl in cfg2 cfg5; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus 2 --steps 10 --warmup 3 --lite --workload $wl > gpurun_out/final_n2_$wl.json 2> gpurun_out/final_n2_$wl.err
echo "bench N=2 $wl rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/final_n2_$wl.json').read().strip().splitlines()[-1])
    print('$wl', {k:d[k] for k in ['value','ms_per_step','phase_ms_sequential']}, (d.get('dp_parity') or {}).get('ok'))
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/final_n2_$wl.err').read()[-2500:])
PY
done
