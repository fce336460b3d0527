#!/bin/bash
# split long-context attention in the megakernel: full suite, bench (headline must not move; config 3 B=1 decode should)
T=${1:-r2sa}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -2 gpurun_out/${T}_bench.err
PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 8 32 > gpurun_out/${T}_decode.txt 2>&1
cat gpurun_out/${T}_decode.txt
