#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2l}
python -m pytest tests/test_gpu_mimi.py tests/test_gpu_generator.py "tests/test_gpu_fullsize.py::test_mimi_decode_60s_batch4_vs_oracle" -m gpu -q -x -s > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
grep -E "passed|failed|rc=|Error|error|SNR|assert" gpurun_out/${T}_tests.log | tail -8
python tools/bench_mimi.py 8 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_mimi_launches.csv python tools/prof_mimi.py > gpurun_out/${T}_ncu_mimi.log 2>&1
