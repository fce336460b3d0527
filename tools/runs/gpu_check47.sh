#!/bin/bash
# short prompts (9 .. 63 rows) through the row-batched path (skinny / tcgen05 GEMMs) instead of the per-op small-row passes
T=${1:-r2pf}
mkdir -p gpurun_out
CSM_PREFILL_TC_MIN=9 timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -6 gpurun_out/${T}_tests.log
for k in 9 64; do
  CSM_PREFILL_TC_MIN=$k timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_min$k.json 2> gpurun_out/${T}_bench$k.err
done
