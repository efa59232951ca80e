#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/lds_bench > gpurun_out/lds_bench.log 2>&1; echo "lds rc=$?"
timeout -k 10 300 python tools/rollout_trace.py > gpurun_out/rollout_trace.log 2>&1; echo "trace rc=$?"; tail -60 gpurun_out/rollout_trace.log
SB200_CUDA_GRAPH=0 CUDA_LAUNCH_BLOCKING=1 timeout -k 10 300 python -m pytest "tests/test_epoch_kernel_gpu.py::test_epoch_kernel_matches_launch_chain[4]" "tests/test_epoch_kernel_gpu.py::test_epoch_kernel_matches_launch_chain[5]" -m gpu -q -x -s --timeout=300 -p no:cacheprovider > gpurun_out/pytest_case4.log 2>&1
echo "case4 eager rc=$?"; grep -n "Error\|error\|ops.py\|ppo.py" gpurun_out/pytest_case4.log | head -30; tail -5 gpurun_out/pytest_case4.log
