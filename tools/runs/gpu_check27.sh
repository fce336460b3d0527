#!/bin/bash
# cluster split-K, reduce loads in flight: unit tests, step times for two thresholds, per-shape profile
T=${1:-r2ck2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q -m gpu > gpurun_out/${T}_unit.log 2>&1
tail -3 gpurun_out/${T}_unit.log
if ! grep -q "passed" gpurun_out/${T}_unit.log || grep -q "failed" gpurun_out/${T}_unit.log; then exit 1; fi
for kb in 32 16; do
  echo "== CSM_TC_CLUSTER_MIN_KB=$kb" >> gpurun_out/${T}_decode.txt
  CSM_TC_CLUSTER_MIN_KB=$kb PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 40 64 128 256 >> gpurun_out/${T}_decode.txt 2>&1
done
cat gpurun_out/${T}_decode.txt
PF_B=256 PF_SHORT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_decode_batch.py > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
