#!/bin/bash
# Round-2 GPU check: tests, smoke, fence on/off bench, sanitizer on the tiny smoke.  Run under gpurun.
set -u
mkdir -p gpurun_out
T=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q -s > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
for rep in 1 2; do
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_fence_$rep.json 2> gpurun_out/${T}_bench_fence_$rep.err
  CSM_B200_LIB=$PWD/sesameai-tts_b200/lib/libcsm_b200_nofence.so python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_nofence_$rep.json 2> gpurun_out/${T}_bench_nofence_$rep.err
done
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${T}_racecheck.log
tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_bench_*.json | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
    except Exception as e: print('bad', e)
"
tail -3 gpurun_out/${T}_memcheck.log; tail -3 gpurun_out/${T}_racecheck.log
