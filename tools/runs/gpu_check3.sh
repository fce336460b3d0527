#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2c}
python -m pytest tests/test_gpu_post.py tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
for rep in 1 2; do
for v in f0 f1 f2 f4 f7; do
  lib=$PWD/sesameai-tts_b200/lib/libcsm_b200_$v.so
  [ "$v" = f7 ] && lib=$PWD/sesameai-tts_b200/lib/libcsm_b200.so
  CSM_B200_LIB=$lib python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_${v}_$rep.json 2> gpurun_out/${T}_bench_${v}_$rep.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${T}_bench_${v}_$rep.json').read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['e2e']['value'])"
done
done
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --fast > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err
tail -c 3000 gpurun_out/${T}_bench_full.json
tail -5 gpurun_out/${T}_bench_full.err
grep -E "passed|failed|rc=" gpurun_out/${T}_tests.log | tail -3
