#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_rollout_oracle_gpu.py tests/test_replay_rollout_gpu.py tests/test_kernels_gpu.py tests/test_ppo_learner_gpu.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/pytest_rollout.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/pytest_rollout.log
timeout -k 10 120 python tools/rollout_trace.py > gpurun_out/rollout_trace_v2b.log 2>&1; echo "trace rc=$?"; head -18 gpurun_out/rollout_trace_v2b.log
bl() { # tag, env...
  env $2 $3 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --lite > gpurun_out/bench_lite_$1.json 2> gpurun_out/bench_lite_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_lite_$1.json').read().strip().splitlines()[-1])
    print('$1', {k:d[k] for k in ['value','ms_per_step','phase_ms_sequential']}, d['kernel_breakdown'][0]['avg_us'])
except Exception as e:
    print('no bench line $1', e); print(open('gpurun_out/bench_lite_$1.err').read()[-1500:])
PY
}
bl default X=1
bl head0 SB200_RF_HEAD=0
bl l1split SB200_RF_L1DIRECT=0
timeout -k 10 120 python tools/prof_gae.py > gpurun_out/prof_gae.log 2>&1; tail -1 gpurun_out/prof_gae.log
SB200_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --lite --sequential > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ppo_rollout2_kernel -s 2 -c 1 \
    -o gpurun_out/prof_rollout -f python tools/prof_rollout.py 128 1 > gpurun_out/prof_rollout.log 2>&1
echo "rollout capture rc=$?"
