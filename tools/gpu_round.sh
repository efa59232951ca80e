#!/bin/bash
# One GPU-box visit: tcgen05 harness first (cheap, no Python); the python legs only enable the tcgen05 path when it passed.
# usage: tools/gpu_round.sh [all|harness|prof|tests|bench] ...
mkdir -p gpurun_out
what=" ${*:-all} "
has() { [[ "$what" == *" all "* || "$what" == *" $1 "* ]]; }
timeout 180 tools/tc5_harness > gpurun_out/tc5_harness.log 2>&1
echo "harness rc=$?" >> gpurun_out/tc5_harness.log
if grep -q "HARNESS OK" gpurun_out/tc5_harness.log; then export SB200_TC5=1; else export SB200_TC5=0; fi
echo "SB200_TC5=$SB200_TC5" >> gpurun_out/tc5_harness.log
tail -14 gpurun_out/tc5_harness.log
if has prof; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp3_tc5 -s 2 -c 1 -f -o gpurun_out/prof_tc5 tools/tc5_harness prof > gpurun_out/prof_tc5.log 2>&1
  echo "ncu rc=$?"; tail -3 gpurun_out/prof_tc5.log
fi
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -30 gpurun_out/pytest_gpu.log
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
fi
if has trace; then
  timeout 120 tools/tc5_trace trace > gpurun_out/tc5_trace.log 2>&1
  echo "trace rc=$?"; tail -34 gpurun_out/tc5_trace.log
fi
if has newtests; then
  timeout 900 python -m pytest tests/test_stem_gpu.py tests/test_checkpoint_gpu.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1
  echo "pytest-new rc=$?" >> gpurun_out/pytest_new.log
  tail -30 gpurun_out/pytest_new.log
fi
