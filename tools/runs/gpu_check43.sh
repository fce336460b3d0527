#!/bin/bash
# once-per-codebook-step KV ordering (MEGA_KV_FENCE=8, mega::kv_step_sync) against the default build: headline bench
# twice each, alternating, then the megakernel parity / stress tests on the fenced library
T=${1:-r2f8}
mkdir -p gpurun_out
F8=$PWD/sesameai-tts_b200/lib/libcsm_b200_f8.so
for rep in 1 2; do
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_default_$rep.json 2> gpurun_out/${T}_bench_default_$rep.err
  CSM_B200_LIB=$F8 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_f8_$rep.json 2> gpurun_out/${T}_bench_f8_$rep.err
done
T=$T python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/'+'%s'%__import__('os').environ.get('T','r2f8')+'_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
CSM_B200_LIB=$F8 timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_tests.log
