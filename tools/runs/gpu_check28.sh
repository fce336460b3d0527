#!/bin/bash
# depth-decoder attention kernel + everything of today: full suite, decode step times, bench
T=${1:-r2ad}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -5 gpurun_out/${T}_tests.log
PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 2 8 32 64 128 256 > gpurun_out/${T}_decode.txt 2>&1
cat gpurun_out/${T}_decode.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -2 gpurun_out/${T}_bench.err
