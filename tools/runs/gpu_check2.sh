#!/bin/bash
# tests + fence-variant timing
set -u
mkdir -p gpurun_out
T=${1:-r2b}
python -m pytest tests -m gpu -q -s > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
for v in f0 f1 f2 f4 f7; do
  lib=$PWD/sesameai-tts_b200/lib/libcsm_b200_$v.so
  [ "$v" = f7 ] && lib=$PWD/sesameai-tts_b200/lib/libcsm_b200.so
  CSM_B200_LIB=$lib python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${T}_bench_$v.json').read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['e2e']['value'])"
done
grep -E "passed|failed" gpurun_out/${T}_tests.log | tail -3
