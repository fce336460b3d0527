#!/bin/bash
# final tree of round 2 (KV rows released / acquired once per codebook step by default): what the driver runs at round
# end (GPU tests, smoke, both bench arms), launch list + one ncu --set full capture of the megakernel (+ one capture
# restricted by NVTX range), and the headline of the unfenced (-DMEGA_KV_FENCE=0) and inline (=24) builds on the same box
set -u
mkdir -p gpurun_out
T=${1:-r2fin6}
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
grep -E "passed|failed|rc=|real" gpurun_out/${T}_tests.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_smoke.log
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
tail -c 300 gpurun_out/${T}_bench_reference.json
( time python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -4 gpurun_out/${T}_bench.err; head -c 700 gpurun_out/${T}_bench.json; echo
CSM_B200_LIB=$PWD/sesameai-tts_b200/lib/libcsm_b200_f0.so timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_unfenced.json 2> gpurun_out/${T}_bench_unfenced.err
CSM_B200_LIB=$PWD/sesameai-tts_b200/lib/libcsm_b200_f24.so timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_inline.json 2> gpurun_out/${T}_bench_inline.err
timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_headline.json 2> gpurun_out/${T}_bench_headline.err
head -c 200 gpurun_out/${T}_bench_inline.json; echo
head -c 200 gpurun_out/${T}_bench_unfenced.json; echo; head -c 200 gpurun_out/${T}_bench_headline.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'k_frame_mega|k_mega_prepare|k_gemv|k_skinny|k_embed|k_attn|k_sample|k_set_|k_rmsnorm|k_rope|k_gemm' -s 100 -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame_mega -s 10 -c 1 -o gpurun_out/${T}_mega python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_ncu_mega.log 2>&1
ls -la gpurun_out/${T}_mega.ncu-rep
timeout 200 ncu --nvtx --nvtx-include "csm.decode.mega/" --metrics gpu__time_duration.sum --clock-control none -c 6 --csv --log-file gpurun_out/${T}_nvtx_mega_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_ncu_nvtx.log 2>&1
tail -8 gpurun_out/${T}_nvtx_mega_launches.csv | cut -c1-220
