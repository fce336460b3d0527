#!/bin/bash
# fused depth-decoder attention: softmax rows dealt over all eight warps + P.V tiles unrolled (MEGA_PV_UNROLL 4 / 2 / 1)
# against the tree before (base); headline bench twice each, alternating; parity / stress tests on the default build
T=${1:-r2pv}
mkdir -p gpurun_out
L=$PWD/sesameai-tts_b200/lib
for rep in 1 2; do
  for v in base pv4 pv2 pv1; do
    if [ $v = pv4 ]; then unset CSM_B200_LIB; else export CSM_B200_LIB=$L/libcsm_b200_$v.so; fi
    timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_${v}_$rep.json 2> gpurun_out/${T}_bench_${v}_$rep.err
  done
done
unset CSM_B200_LIB
T=$T python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/%s_bench_*.json' % os.environ['T'])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'], 4), round(d['value'], 1), round(d['e2e']['value'], 1))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 120 python tools/trace_mega.py > gpurun_out/${T}_trace_pv4.txt 2>&1; tail -12 gpurun_out/${T}_trace_pv4.txt
timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_tests.log
