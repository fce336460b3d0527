#!/bin/bash
# tall-tile gate/up GEMM + skinny limit 32: unit tests, full suite, decode step times with / without the tall CTA
T=${1:-r2tall}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q -m gpu > gpurun_out/${T}_unit.log 2>&1
tail -3 gpurun_out/${T}_unit.log
if ! grep -q "passed" gpurun_out/${T}_unit.log || grep -q "failed" gpurun_out/${T}_unit.log; then exit 1; fi
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
for k in 1 0; do
  echo "== CSM_TC_TALL=$k" >> gpurun_out/${T}_decode.txt
  CSM_TC_TALL=$k PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 64 128 200 256 >> gpurun_out/${T}_decode.txt 2>&1
done
cat gpurun_out/${T}_decode.txt
