#!/bin/bash
# ncu --set full of skinny launches inside a 32-stream decode step (decoder layers: qkv / o / gate-up / down)
T=${1:-r2sk32}
mkdir -p gpurun_out
PF_B=32 PF_SHORT=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_skinny -s 80 -c 8 -o gpurun_out/${T}_skinny python tools/prof_decode_batch.py > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
