#!/bin/bash
# Run on the GPU box (gpurun): launch list of one bench step + one full ncu capture of the dominant kernel.
# Outputs land in gpurun_out/ ; summaries are copied into profiles/ by tools/summarize_profiles.py here.
set -x
mkdir -p gpurun_out
export SB200_CUDA_GRAPH=${SB200_CUDA_GRAPH:-0}      # eager launches: every kernel is a separate ncu record
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2450 -c 850 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --lite > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:mlp_fwd_kernelILi8E -s 3 -c 1 \
    -o gpurun_out/prof_critic -f python bench.py --steps 1 --warmup 3 --lite > gpurun_out/prof_critic.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gae_full -c 1 -o gpurun_out/prof_gae -f \
    python bench.py --steps 1 --warmup 3 --lite > gpurun_out/prof_gae.log 2>&1
ls -la gpurun_out
