#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2i}
python -m pytest tests/test_gpu_frame.py tests/test_gpu_fullsize.py tests/test_gpu_serving.py tests/test_gpu_generator.py tests/test_checkpoint_roundtrip.py -m gpu -q -x -s > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
grep -E "passed|failed|rc=|Error|service loop" gpurun_out/${T}_tests.log | tail -8
for rep in 1 2; do
(cd r1_tree && python bench.py --steps 100 --warmup 5 --no-cpu-baseline) > gpurun_out/${T}_bench_r1tree_$rep.json 2> gpurun_out/${T}_bench_r1tree_$rep.err
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_cur_$rep.json 2> gpurun_out/${T}_bench_cur_$rep.err
CSM_MEGA_NO_KINT=1 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_nokint_$rep.json 2> gpurun_out/${T}_bench_nokint_$rep.err
for f in r1tree cur nokint; do python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_${f}_$rep.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['e2e']['value'])"; done
done
python tools/trace_mega.py > gpurun_out/${T}_trace.log 2>&1; tail -30 gpurun_out/${T}_trace.log
