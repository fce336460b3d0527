#!/bin/bash
# KV ordering once per codebook step (default build, MEGA_KV_FENCE=24): release after gate/up + acquire before sampling
# (place 0) vs release before sampling + acquire before the next step's first phase (CSM_MEGA_KV_PLACE=1) vs the
# unfenced build (-DMEGA_KV_FENCE=0); headline bench twice each, alternating; parity / stress tests on place 1
T=${1:-r2place}
mkdir -p gpurun_out
L=$PWD/sesameai-tts_b200/lib
for rep in 1 2; do
  timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_place0_$rep.json 2> gpurun_out/${T}_bench_place0_$rep.err
  CSM_MEGA_KV_PLACE=1 timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_place1_$rep.json 2> gpurun_out/${T}_bench_place1_$rep.err
  CSM_B200_LIB=$L/libcsm_b200_f0.so timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_unfenced_$rep.json 2> gpurun_out/${T}_bench_unfenced_$rep.err
done
T=$T python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/%s_bench_*.json' % os.environ['T'])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'], 4), round(d['value'], 1), round(d['e2e']['value'], 1), d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
CSM_MEGA_KV_PLACE=1 timeout 120 python tools/trace_mega.py > gpurun_out/${T}_trace_place1.txt 2>&1; head -16 gpurun_out/${T}_trace_place1.txt
CSM_MEGA_KV_PLACE=1 timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_tests.log
