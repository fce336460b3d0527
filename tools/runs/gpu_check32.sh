#!/bin/bash
# RoPE + KV append in the cluster split-K epilogue: full suite, decode step times with / without
T=${1:-r2rope}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
for k in 1 0; do
  echo "== CSM_TC_ROPE_FUSE=$k" >> gpurun_out/${T}_decode.txt
  CSM_TC_ROPE_FUSE=$k PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 64 128 256 >> gpurun_out/${T}_decode.txt 2>&1
done
cat gpurun_out/${T}_decode.txt
