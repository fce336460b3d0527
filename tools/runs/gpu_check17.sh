#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2s}
PF_SHORT=1 python tools/bench_decode_batch.py 16 32 2>&1 | tail -2
CSM_SKINNY_MAX_ROWS=8 PF_SHORT=1 python tools/bench_decode_batch.py 16 32 2>&1 | tail -2
CSM_SKINNY_MAX_ROWS=8 CSM_TC_SPLITK=1 PF_SHORT=1 python tools/bench_decode_batch.py 16 32 2>&1 | tail -2
