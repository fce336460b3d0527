#!/bin/bash
# k_attn_gqa64 (backbone decode attention, one CTA per (row, KV head)): full suite, config 3 decode with / without, step times
T=${1:-r2gqa}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
for k in 1 0; do
  CSM_ATTN_GQA=$k timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_gqa$k.json 2> gpurun_out/${T}_bench$k.err
  echo "== CSM_ATTN_GQA=$k" >> gpurun_out/${T}_decode.txt
  CSM_ATTN_GQA=$k PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 8 32 256 >> gpurun_out/${T}_decode.txt 2>&1
done
cat gpurun_out/${T}_decode.txt
