#!/bin/bash
# 2+-GPU visit: DP parity tests (peer all-reduce path and NCCL path) and a short N-GPU bench
mkdir -p gpurun_out
N=${1:-2}
timeout 120 tools/tc5_harness prof > gpurun_out/tc5_prof.log 2>&1; tail -2 gpurun_out/tc5_prof.log
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1
echo "dp rc=$?"; tail -15 gpurun_out/pytest_dp.log
SB200_PEER_ALLREDUCE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/dp_check.py > gpurun_out/dp_nccl.log 2>&1
echo "dp-nccl rc=$?"; grep -E "DP_OK|DP_FAIL" gpurun_out/dp_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','phase_ms_sequential','dp_parity']})
    print('e2e', d['e2e']['value'])
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/bench_n$N.err').read()[-3000:])
PY
SB200_PEER_ALLREDUCE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus $N --steps 10 --warmup 3 --lite > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_nccl.json').read().strip().splitlines()[-1])
    print('NCCL path:', {k:d[k] for k in ['value','ms_per_step','phase_ms_sequential']})
except Exception as e:
    print('no nccl bench line', e)
PY
