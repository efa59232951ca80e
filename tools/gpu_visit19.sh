#!/bin/bash
mkdir -p gpurun_out
export SB200_RF_ASYNC=1
timeout -k 10 600 python -m pytest tests/test_rollout_oracle_gpu.py tests/test_replay_rollout_gpu.py tests/test_kernels_gpu.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/pytest_rollout.log 2>&1
echo "rollout tests (async) rc=$?"; tail -5 gpurun_out/pytest_rollout.log
timeout -k 10 120 python tools/rollout_trace.py > gpurun_out/rollout_trace_async.log 2>&1; echo "trace rc=$?"; head -18 gpurun_out/rollout_trace_async.log
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --lite > gpurun_out/bench_lite_async.json 2> gpurun_out/bench_lite_async.err
echo "bench async rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_lite_async.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','phase_ms_sequential']}); print(d['kernel_breakdown'][0]); print(d['extras'] and d['extras']['gae_sweep']['cases'][-1])
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/bench_lite_async.err').read()[-2000:])
PY
timeout -k 10 120 python tools/prof_gae.py > gpurun_out/prof_gae.log 2>&1; tail -1 gpurun_out/prof_gae.log
