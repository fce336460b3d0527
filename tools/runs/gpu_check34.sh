#!/bin/bash
# 128 x 256 tiles in the persistent GEMM: unit tests, full suite, bench (config 3 prefill) with 256- and 128-wide tiles,
# 256-stream decode step both ways
T=${1:-r2wide}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q -m gpu > gpurun_out/${T}_unit.log 2>&1
tail -3 gpurun_out/${T}_unit.log
if ! grep -q "passed" gpurun_out/${T}_unit.log || grep -q "failed" gpurun_out/${T}_unit.log; then exit 1; fi
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
for bn in 256 128; do
  CSM_TC_BN=$bn timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_bn$bn.json 2> gpurun_out/${T}_bench$bn.err
  echo "== CSM_TC_BN=$bn" >> gpurun_out/${T}_decode.txt
  CSM_TC_BN=$bn PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 128 256 >> gpurun_out/${T}_decode.txt 2>&1
done
cat gpurun_out/${T}_decode.txt
