#!/bin/bash
# warp-per-row RMSNorm, skinny limit 32: full suite, decode step times
T=${1:-r2rn}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 2 8 16 32 64 128 256 > gpurun_out/${T}_decode.txt 2>&1
cat gpurun_out/${T}_decode.txt
