#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2p}
PF_B=8 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_prefill_B8_launches.csv python tools/prof_prefill.py > gpurun_out/${T}_ncu_prefill.log 2>&1
PF_SHORT=1 python tools/bench_decode_batch.py 64 128 256 > gpurun_out/${T}_decode_batch_short.log 2>&1
tail -3 gpurun_out/${T}_decode_batch_short.log
CSM_TC_ONE_TILE=1 PF_SHORT=1 python tools/bench_decode_batch.py 256 2>&1 | tail -1
