#!/bin/bash
# RoPE + KV append in the persistent GEMM epilogue (prefill): tests + config 3 prefill with / without
T=${1:-r2rope2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
for k in 1 0; do
  CSM_TC_ROPE_FUSE=$k timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_fuse$k.json 2> gpurun_out/${T}_bench$k.err
done
