#!/bin/bash
# rollout v2 validation: parity tests, trace, bench (v2 and v1)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_rollout_oracle_gpu.py tests/test_replay_rollout_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_rollout.log 2>&1
echo "rollout tests rc=$?"; tail -6 gpurun_out/pytest_rollout.log
timeout -k 10 300 python tools/rollout_trace.py > gpurun_out/rollout_trace_v2.log 2>&1; echo "trace rc=$?"; head -20 gpurun_out/rollout_trace_v2.log
for v in 1 0; do
SB200_RF_V2=$v SB200_EPOCH_KERNEL=0 timeout -k 10 600 python bench.py --steps 20 --warmup 5 --lite > gpurun_out/bench_lite_v$v.json 2> gpurun_out/bench_lite_v$v.err
echo "bench v2=$v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_lite_v$v.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','phase_ms_sequential','gpu_launches_per_step']})
    for k in d['kernel_breakdown'][:4]: print(k)
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/bench_lite_v$v.err').read()[-2000:])
PY
done
