#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2h}
python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
grep -E "passed|failed|rc=|Error" gpurun_out/${T}_tests.log | tail -3
PF_SHORT=1 python tools/bench_decode_batch.py 64 128 256 > gpurun_out/${T}_decode_batch_short.log 2>&1
tail -3 gpurun_out/${T}_decode_batch_short.log
for rep in 1 2; do
(cd r1_tree && python bench.py --steps 100 --warmup 5 --no-cpu-baseline) > gpurun_out/${T}_bench_r1tree_$rep.json 2> gpurun_out/${T}_bench_r1tree_$rep.err
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_cur_$rep.json 2> gpurun_out/${T}_bench_cur_$rep.err
for v in v1 v2 v3; do
CSM_B200_LIB=$PWD/sesameai-tts_b200/lib/libcsm_b200_$v.so python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_${v}_$rep.json 2> gpurun_out/${T}_bench_${v}_$rep.err
done
for f in r1tree cur v1 v2 v3; do python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_${f}_$rep.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['e2e']['value'])"; done
done
