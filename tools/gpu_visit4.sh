#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/prof_epoch2.py > gpurun_out/prof_epoch2.log 2>&1; echo "prof rc=$?"; tail -16 gpurun_out/prof_epoch2.log
timeout -k 10 900 python -m pytest tests/test_epoch_kernel_gpu.py tests/test_rollout_oracle_gpu.py tests/test_replay_rollout_gpu.py tests/test_ppo_learner_gpu.py tests/test_fullsize_gpu.py tests/test_kernels_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_a.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/pytest_a.log
timeout -k 10 300 python tools/rollout_trace.py > gpurun_out/rollout_trace_v3.log 2>&1; echo "trace rc=$?"; head -18 gpurun_out/rollout_trace_v3.log
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --lite > gpurun_out/bench_lite.json 2> gpurun_out/bench_lite.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_lite.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','phase_ms_sequential','gpu_launches_per_step']})
    for k in d['kernel_breakdown'][:8]: print(k)
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/bench_lite.err').read()[-2000:])
PY
