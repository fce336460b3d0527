#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2k}
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_mimi_launches.csv python tools/prof_mimi.py > gpurun_out/${T}_ncu_mimi.log 2>&1
ls -la gpurun_out/${T}_mimi_launches.csv
