#!/bin/bash
# prefill after the producer-loop fix; skinny vs tcgen05 crossover with cluster split-K
T=${1:-r2x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q -m gpu > gpurun_out/${T}_unit.log 2>&1
tail -2 gpurun_out/${T}_unit.log
for m in 64 32 16 8; do
  echo "== CSM_SKINNY_MAX_ROWS=$m" >> gpurun_out/${T}_decode.txt
  CSM_SKINNY_MAX_ROWS=$m PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 8 16 24 32 48 64 >> gpurun_out/${T}_decode.txt 2>&1
done
cat gpurun_out/${T}_decode.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -2 gpurun_out/${T}_bench.err
