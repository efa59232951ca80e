#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 120 python tools/rollout_trace.py > gpurun_out/rollout_trace_v2d.log 2>&1; echo "trace rc=$?"; head -19 gpurun_out/rollout_trace_v2d.log
