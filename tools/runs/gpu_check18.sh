#!/bin/bash
# programmatic dependent launch in the row-batched path: parity tests, then decode timings with and without it
T=${1:-r2pdl}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
tail -8 gpurun_out/${T}_tests.log
for pdl in 1 0 1 0; do
  echo "== CSM_PDL=$pdl" >> gpurun_out/${T}_decode.txt
  CSM_PDL=$pdl PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 2 8 32 64 256 >> gpurun_out/${T}_decode.txt 2>&1
done
cat gpurun_out/${T}_decode.txt
