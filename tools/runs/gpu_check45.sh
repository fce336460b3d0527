#!/bin/bash
# tensor-core flash-decoding attention (k_attn_split64_mma): full suite, config 3 decode vs the CUDA-core kernel
T=${1:-r2am}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
for k in 1 0; do
  CSM_ATT_MMA=$k timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_mma$k.json 2> gpurun_out/${T}_bench$k.err
done
