#!/bin/bash
# full GPU suite on the current build; skinny fused-norm row threshold; Mimi launch list with the tap-shift operand
T=${1:-r2n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -5 gpurun_out/${T}_tests.log
for n in 16 32; do
  echo "== CSM_SKINNY_NORM_ROWS=$n" >> gpurun_out/${T}_decode.txt
  CSM_SKINNY_NORM_ROWS=$n PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 20 24 32 >> gpurun_out/${T}_decode.txt 2>&1
done
cat gpurun_out/${T}_decode.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_mimi_launches.csv python tools/prof_mimi.py > gpurun_out/${T}_ncu_mimi.log 2>&1
tail -2 gpurun_out/${T}_ncu_mimi.log
