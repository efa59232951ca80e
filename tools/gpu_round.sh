#!/bin/bash
# One GPU-box visit: tcgen05 harness first (cheap, no Python); the python legs only enable the tcgen05 path when it passed.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 180 tools/tc5_harness > gpurun_out/tc5_harness.log 2>&1
echo "harness rc=$?" >> gpurun_out/tc5_harness.log
if grep -q "HARNESS OK" gpurun_out/tc5_harness.log; then export SB200_TC5=1; else export SB200_TC5=0; fi
echo "SB200_TC5=$SB200_TC5" >> gpurun_out/tc5_harness.log
tail -20 gpurun_out/tc5_harness.log
if [ "${1:-all}" = "harness" ]; then exit 0; fi
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
