#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_rollout_oracle_gpu.py tests/test_replay_rollout_gpu.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/pytest_rollout.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/pytest_rollout.log
timeout -k 10 120 python tools/rollout_trace.py > gpurun_out/rollout_trace_v2c.log 2>&1; echo "trace rc=$?"; head -18 gpurun_out/rollout_trace_v2c.log
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --lite > gpurun_out/bench_lite.json 2> gpurun_out/bench_lite.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_lite.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','phase_ms_sequential']}, d['kernel_breakdown'][0]['avg_us'])
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/bench_lite.err').read()[-1500:])
PY
