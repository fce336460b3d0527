#!/bin/bash
# continuous batcher with device-side bookkeeping: serving tests, config 5 twice
T=${1:-r2srv}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "serving or generator" > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log
for i in 1 2; do
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench$i.json 2> gpurun_out/${T}_bench$i.err
done
