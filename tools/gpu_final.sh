#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, both bench arms (+ a launch list that reaches the decode frames)
set -u
mkdir -p gpurun_out
T=${1:-r2final}
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
grep -E "passed|failed|rc=|real" gpurun_out/${T}_tests.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_smoke.log
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
tail -c 600 gpurun_out/${T}_bench_reference.json; tail -4 gpurun_out/${T}_bench_reference.err
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -4 gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'k_frame_mega|k_mega_prepare|k_gemv|k_skinny|k_embed|k_attn|k_sample|k_set_|k_rmsnorm|k_rope|k_gemm' -s 100 -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32 -s 60 -c 3 -o gpurun_out/${T}_mimi_gemm python tools/prof_mimi.py > gpurun_out/${T}_ncu_mimi_gemm.log 2>&1
