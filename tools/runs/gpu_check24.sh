#!/bin/bash
# per-shape GEMM times of one 256-stream decode step (direct launches under ncu)
T=${1:-r2b256}
mkdir -p gpurun_out
PF_B=256 PF_SHORT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_decode_batch.py > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
