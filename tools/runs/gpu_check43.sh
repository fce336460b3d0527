#!/bin/bash
# flash-decoding attention for row-batched decode at long context: full suite, config 3 decode with / without
T=${1:-r2fd}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
for k in 256 0; do
  CSM_ATT_LONG_MIN=$k timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_long$k.json 2> gpurun_out/${T}_bench$k.err
done
