#!/bin/bash
# epoch2 validation: new-vs-chain, goldens, full-size oracle cases, bench
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_epoch_kernel_gpu.py -m gpu -q -x -s --timeout=300 -p no:cacheprovider > gpurun_out/pytest_epoch.log 2>&1
echo "epoch rc=$?"; tail -25 gpurun_out/pytest_epoch.log
timeout -k 10 900 python -m pytest tests/test_ppo_learner_gpu.py tests/test_fullsize_gpu.py tests/test_checkpoint_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/pytest_ppo.log 2>&1
echo "ppo rc=$?"; tail -12 gpurun_out/pytest_ppo.log
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
