#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/engine_timeline.py > gpurun_out/engine_timeline.log 2>&1; echo "timeline rc=$?"; tail -8 gpurun_out/engine_timeline.log
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --lite > gpurun_out/bench_lite.json 2> gpurun_out/bench_lite.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_lite.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','phase_ms_sequential','gpu_launches_per_step']})
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/bench_lite.err').read()[-2000:])
PY
