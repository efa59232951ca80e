#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/fma_bench > gpurun_out/fma_bench.log 2>&1; cat gpurun_out/fma_bench.log
timeout -k 10 300 python tools/prof_epoch2.py > gpurun_out/prof_epoch2.log 2>&1; echo "prof FFMA2 rc=$?"; tail -44 gpurun_out/prof_epoch2.log | head -22
SB200_LIB=$PWD/tools/_trace/libsurreal_b200_scalar.so timeout -k 10 300 python tools/prof_epoch2.py > gpurun_out/prof_epoch2_scalar.log 2>&1; echo "prof scalar rc=$?"; tail -44 gpurun_out/prof_epoch2_scalar.log | head -22
timeout -k 10 600 python -m pytest tests/test_epoch_kernel_gpu.py tests/test_ppo_learner_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_a.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/pytest_a.log
