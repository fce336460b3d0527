#!/bin/bash
# once-per-codebook-step KV ordering: fence + relaxed (MEGA_KV_FENCE=8) vs st.release / ld.acquire (=24) vs the default
# build; headline bench twice each, alternating; per-phase trace of the default and the =24 kernel; parity tests on =24
T=${1:-r2f24}
mkdir -p gpurun_out
L=$PWD/sesameai-tts_b200/lib
for rep in 1 2; do
  for v in default f8 f24; do
    if [ $v = default ]; then unset CSM_B200_LIB; else export CSM_B200_LIB=$L/libcsm_b200_$v.so; fi
    python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_${v}_$rep.json 2> gpurun_out/${T}_bench_${v}_$rep.err
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
python tools/trace_mega.py > gpurun_out/${T}_trace_default.txt 2>&1
CSM_B200_LIB=$L/libcsm_b200_f24.so python tools/trace_mega.py > gpurun_out/${T}_trace_f24.txt 2>&1
head -16 gpurun_out/${T}_trace_default.txt; head -16 gpurun_out/${T}_trace_f24.txt
CSM_B200_LIB=$L/libcsm_b200_f24.so timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_tests.log
