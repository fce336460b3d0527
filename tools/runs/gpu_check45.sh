#!/bin/bash
# KV ordering agent warp (MEGA_KV_FENCE=32) vs inline st.release / ld.acquire (=24) vs the default build; headline bench
# twice each, alternating; per-phase trace and parity / stress tests on the agent build
T=${1:-r2f32}
mkdir -p gpurun_out
L=$PWD/sesameai-tts_b200/lib
for rep in 1 2; do
  for v in default f24 f32; do
    if [ $v = default ]; then unset CSM_B200_LIB; else export CSM_B200_LIB=$L/libcsm_b200_$v.so; fi
    timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_${v}_$rep.json 2> gpurun_out/${T}_bench_${v}_$rep.err
  done
done
unset CSM_B200_LIB
T=$T python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/%s_bench_*.json' % os.environ['T'])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'], 4), round(d['value'], 1), round(d['e2e']['value'], 1), d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
CSM_B200_LIB=$L/libcsm_b200_f32.so timeout 120 python tools/trace_mega.py > gpurun_out/${T}_trace_f32.txt 2>&1; head -16 gpurun_out/${T}_trace_f32.txt
CSM_B200_LIB=$L/libcsm_b200_f32.so timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_tests.log
