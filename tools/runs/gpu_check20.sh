#!/bin/bash
# batched Mimi transformer + resident tail: tests, timings (B=8 / B=64 config 4), ncu of the fused tail GEMM
T=${1:-r2bt}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "mimi or generator or post or dependent_launch" > gpurun_out/${T}_tests.log 2>&1
tail -8 gpurun_out/${T}_tests.log
for r in 1 0; do
  echo "== MIMI_BATCH=$r" >> gpurun_out/${T}_mimi.txt
  MIMI_BATCH=$r timeout 300 python tools/bench_mimi.py 8 >> gpurun_out/${T}_mimi.txt 2>&1
  MIMI_BATCH=$r timeout 300 python tools/bench_mimi.py 64 >> gpurun_out/${T}_mimi.txt 2>&1
done
cat gpurun_out/${T}_mimi.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32_r -s 6 -c 2 -o gpurun_out/${T}_tail python tools/prof_mimi.py > gpurun_out/${T}_ncu_tail.log 2>&1
tail -3 gpurun_out/${T}_ncu_tail.log
