#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2d}
# same-box comparison with the round-1 tree
(cd r1_tree && python bench.py --steps 100 --warmup 5 --no-cpu-baseline) > gpurun_out/${T}_bench_r1tree.json 2> gpurun_out/${T}_bench_r1tree.err
CSM_B200_LIB=$PWD/sesameai-tts_b200/lib/libcsm_b200_f0.so python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_f0.json 2> gpurun_out/${T}_bench_f0.err
for f in r1tree f0; do python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['e2e']['value'])"; done
PF_SHORT=1 python tools/bench_decode_batch.py 8 32 64 128 256 > gpurun_out/${T}_decode_batch_short.log 2>&1
cat gpurun_out/${T}_decode_batch_short.log | tail -6
for B in 32 256; do
PF_SHORT=1 PF_B=$B ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_B$B.csv python tools/prof_decode_batch.py > gpurun_out/${T}_ncu_B$B.log 2>&1
done
ls -la gpurun_out/${T}_launches_B*.csv
