#!/bin/bash
# 2-GPU visit (`gpurun --gpus 2`): data-parallel parity tests (in-kernel NVLink exchange; NCCL path) and N=2 bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1
echo "dp rc=$?"; tail -3 gpurun_out/pytest_dp.log
SB200_PEER_ALLREDUCE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/dp_check.py > gpurun_out/dp_nccl.log 2>&1
echo "dp-nccl rc=$?"; grep -E "DP_OK|DP_FAIL" gpurun_out/dp_nccl.log
bash tools/gpu_scale.sh 2
