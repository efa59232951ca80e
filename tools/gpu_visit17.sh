#!/bin/bash
# full single-GPU validation: engine timeline, whole GPU test suite, smoke, both bench arms
mkdir -p gpurun_out
timeout -k 10 300 python tools/engine_timeline.py > gpurun_out/engine_timeline.log 2>&1; echo "timeline rc=$?"; tail -9 gpurun_out/engine_timeline.log
SB200_EPOCH_KERNEL=0 timeout -k 10 300 python tools/engine_timeline.py > gpurun_out/engine_timeline_chain.log 2>&1; echo "timeline(chain) rc=$?"; tail -8 gpurun_out/engine_timeline_chain.log
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; tail -c 300 gpurun_out/bench_ref.json
