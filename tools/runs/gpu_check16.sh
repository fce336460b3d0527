#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2r}
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'k_frame_mega|k_mega_prepare|k_gemv|k_skinny|k_embed|k_attn|k_sample|k_set_|k_rmsnorm|k_rope|k_gemm|k_pack|k_copy|k_interleave|k_transpose' -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame_mega -s 10 -c 1 -o gpurun_out/${T}_mega python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_ncu_mega.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32 -s 95 -c 3 -o gpurun_out/${T}_mimi_gemm python tools/prof_mimi.py > gpurun_out/${T}_ncu_mimi_gemm.log 2>&1
PF_B=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc_p -s 130 -c 2 -o gpurun_out/${T}_prefill_gemm python tools/prof_prefill.py > gpurun_out/${T}_ncu_prefill_gemm.log 2>&1
ls -la gpurun_out/${T}_*
