#!/bin/bash
T=${1:-r2sa2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "frame or fullsize or stress" > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -2 gpurun_out/${T}_bench.err
