#!/bin/bash
# split-K (existing implementation) on the 256-stream decode step: per-shape times under ncu + warm step time
T=${1:-r2sk}
mkdir -p gpurun_out
CSM_TC_SPLITK=1 PF_B=256 PF_SHORT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_decode_batch.py > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
for k in 1 0; do
  echo "== splitk=$k" >> gpurun_out/${T}_decode.txt
  if [ $k = 1 ]; then export CSM_TC_SPLITK=1; else unset CSM_TC_SPLITK; fi
  PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 64 256 >> gpurun_out/${T}_decode.txt 2>&1
done
cat gpurun_out/${T}_decode.txt
