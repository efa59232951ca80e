#!/bin/bash
# Full single-GPU visit: regression, smoke, bench (both arms), pixel line, ncu evidence.
mkdir -p gpurun_out
timeout 120 tools/tc5_harness > gpurun_out/tc5_harness.log 2>&1; tail -3 gpurun_out/tc5_harness.log
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; tail -c 300 gpurun_out/bench_ref.json
timeout 600 python tools/bench_pixel.py 5 > gpurun_out/bench_pixel.json 2> gpurun_out/bench_pixel.err
echo "pixel rc=$?"; cat gpurun_out/bench_pixel.json; tail -3 gpurun_out/bench_pixel.err
bash tools/profile_gpu.sh 2>&1 | tail -20
