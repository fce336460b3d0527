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
# Mimi: launch list of ONE 60 s decode, and the full metric set of the fused tail GEMM (the 4th k_gemm_tf32_r of the second decode)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_mimi_launches.csv python tools/prof_mimi.py > gpurun_out/${T}_ncu_mimi.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32_r -s 7 -c 1 -o gpurun_out/${T}_mimi_tail python tools/prof_mimi.py > gpurun_out/${T}_ncu_mimi_tail.log 2>&1
# one decode step of 256 streams: launch list by kernel
PF_B=256 PF_SHORT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_decode_B256_launches.csv python tools/prof_decode_batch.py > gpurun_out/${T}_ncu_B256.log 2>&1
PF_B=32 PF_SHORT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_decode_B32_launches.csv python tools/prof_decode_batch.py > gpurun_out/${T}_ncu_B32.log 2>&1
PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 2 8 32 64 128 256 > gpurun_out/${T}_decode.txt 2>&1
cat gpurun_out/${T}_decode.txt
