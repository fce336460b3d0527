#!/bin/bash
# skinny GEMM with two CTAs per SM: batched tests + step times
T=${1:-r2occ}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "serving or frame or fullsize or generator or linear or skinny" > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log
PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 2 4 8 16 24 32 > gpurun_out/${T}_decode.txt 2>&1
cat gpurun_out/${T}_decode.txt
