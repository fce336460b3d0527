#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2j}
python -m pytest tests/test_gpu_mimi.py tests/test_gpu_generator.py "tests/test_gpu_fullsize.py::test_mimi_decode_60s_batch4_vs_oracle" -m gpu -q -x -s > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
grep -E "passed|failed|rc=|Error|error|SNR|assert" gpurun_out/${T}_tests.log | tail -12
python tools/bench_mimi.py 8 2>&1 | tail -3
MIMI_DECODE=mma python tools/bench_mimi.py 8 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_cur.json 2> gpurun_out/${T}_bench_cur.err
(cd r1_tree && python bench.py --steps 100 --warmup 5 --no-cpu-baseline) > gpurun_out/${T}_bench_r1tree.json 2> gpurun_out/${T}_bench_r1tree.err
for f in r1tree cur; do python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['e2e']['value'])"; done
