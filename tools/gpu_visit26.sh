#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_epoch_kernel_gpu.py tests/test_ppo_learner_gpu.py tests/test_fullsize_gpu.py tests/test_checkpoint_gpu.py tests/test_rnn_gpu.py tests/test_pixel_gpu.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_b.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/pytest_b.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --lite > gpurun_out/bench_lite.json 2> gpurun_out/bench_lite.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_lite.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ['value','ms_per_step','phase_ms_sequential','gpu_launches_per_step']})
PY
