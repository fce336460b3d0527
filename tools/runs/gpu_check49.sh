#!/bin/bash
# KV ordering on the producer warp (-DMEGA_KV_FENCE=32, kv_producer_poll) vs inline in the consumers (default, =24);
# headline bench twice each, alternating; parity / stress tests and a phase trace on the producer-warp build
T=${1:-r2prod}
mkdir -p gpurun_out
L=$PWD/sesameai-tts_b200/lib
for rep in 1 2; do
  timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_inline_$rep.json 2> gpurun_out/${T}_bench_inline_$rep.err
  CSM_B200_LIB=$L/libcsm_b200_f56.so timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_producer_$rep.json 2> gpurun_out/${T}_bench_producer_$rep.err
done
T=$T python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/%s_bench_*.json' % os.environ['T'])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'], 4), round(d['value'], 1), round(d['e2e']['value'], 1))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/${T}_bench_producer_1.err
CSM_B200_LIB=$L/libcsm_b200_f56.so timeout 120 python tools/trace_mega.py > gpurun_out/${T}_trace_producer.txt 2>&1; head -16 gpurun_out/${T}_trace_producer.txt | tail -8
CSM_B200_LIB=$L/libcsm_b200_f56.so timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_tests.log
