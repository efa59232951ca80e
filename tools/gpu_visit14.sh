#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppo_epochs2 -s 2 -c 1 -f -o gpurun_out/prof_epochs2 \
    python tools/prof_epoch2.py > gpurun_out/prof_epochs2.log 2>&1
echo "epochs2 capture rc=$?"; ls -la gpurun_out/*.ncu-rep
