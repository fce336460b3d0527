#!/bin/bash
# weight-resident Mimi tail GEMM + fused final conv: parity tests, timings, launch list
T=${1:-r2res}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "mimi or generator or post" > gpurun_out/${T}_tests.log 2>&1
tail -8 gpurun_out/${T}_tests.log
for r in 1 0; do
  echo "== MIMI_RESIDENT=$r" >> gpurun_out/${T}_mimi.txt
  MIMI_RESIDENT=$r timeout 300 python tools/bench_mimi.py 8 >> gpurun_out/${T}_mimi.txt 2>&1
done
cat gpurun_out/${T}_mimi.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_mimi_launches.csv python tools/prof_mimi.py > gpurun_out/${T}_ncu_mimi.log 2>&1
tail -3 gpurun_out/${T}_ncu_mimi.log
