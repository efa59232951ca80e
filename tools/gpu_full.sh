#!/bin/bash
# Full single-GPU visit (gpurun): whole GPU test suite, smoke, both bench arms, engine timeline, pixel line.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; tail -c 200 gpurun_out/bench_ref.json
timeout -k 10 300 python tools/engine_timeline.py > gpurun_out/engine_timeline.log 2>&1; echo "timeline rc=$?"; tail -7 gpurun_out/engine_timeline.log
timeout -k 10 300 python tools/prof_epoch2.py > gpurun_out/prof_epoch2.log 2>&1; echo "epoch profile rc=$?"; head -14 gpurun_out/prof_epoch2.log
timeout 600 python tools/bench_pixel.py 5 > gpurun_out/bench_pixel.json 2> gpurun_out/bench_pixel.err
echo "pixel rc=$?"; cat gpurun_out/bench_pixel.json
