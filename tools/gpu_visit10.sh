#!/bin/bash
# 2-GPU visit: DP parity (epoch2 in-kernel exchange), N=1 and N=2 bench lines
mkdir -p gpurun_out
timeout -k 10 300 python tools/prof_epoch2.py > gpurun_out/prof_epoch2.log 2>&1; echo "prof rc=$?"; tail -44 gpurun_out/prof_epoch2.log | head -14
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1
echo "dp rc=$?"; tail -15 gpurun_out/pytest_dp.log
timeout 600 python bench.py --steps 20 --warmup 5 --lite > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench N=1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 10 --warmup 3 --lite > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench N=2 rc=$?"
python - <<'PY'
import json
for n in (1, 2):
    try:
        d=json.loads(open('gpurun_out/bench_n%d.json' % n).read().strip().splitlines()[-1])
        print(n, {k:d[k] for k in ['value','ms_per_step','phase_ms_sequential','dp_parity']})
    except Exception as e:
        print('no bench line', n, e); print(open('gpurun_out/bench_n%d.err' % n).read()[-3000:])
PY
