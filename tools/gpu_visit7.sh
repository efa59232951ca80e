#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/prof_epoch2.py > gpurun_out/prof_epoch2.log 2>&1; echo "prof rc=$?"; tail -44 gpurun_out/prof_epoch2.log
