#!/bin/bash
# last GPU minutes of round 2: code-size / unroll micro-variants of the final megakernel (out-of-line request handler on the
# producer warp, P.V unroll 1 / 4), headline bench each; then the full GPU test suite + smoke() on the fastest one
T=${1:-r2mv}
mkdir -p gpurun_out
L=$PWD/sesameai-tts_b200/lib
run() {  # variant, repetition
  if [ $1 = base ]; then unset CSM_B200_LIB; else export CSM_B200_LIB=$L/libcsm_b200_$1.so; fi
  timeout 100 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_$1_$2.json 2> gpurun_out/${T}_bench_$1_$2.err
  unset CSM_B200_LIB
}
run base 1; run v1 1; run pv4 1; run v1pv4 1; run pv1 1; run base 2; run v1 2
BEST=$(T=$T python - <<'PY'
import json, glob, os, collections
r = collections.defaultdict(list)
for f in sorted(glob.glob('gpurun_out/%s_bench_*.json' % os.environ['T'])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r[f.split('_bench_')[1].rsplit('_', 1)[0]].append(d['ms_per_step'])
    except Exception as e:
        pass
import sys
for k, v in r.items():
    sys.stderr.write('%s %s\n' % (k, ' '.join('%.4f' % x for x in v)))
best = min(r, key=lambda k: sum(r[k]) / len(r[k]))
# a variant has to beat the base by 0.2 % to be worth a change
if 'base' in r and sum(r[best]) / len(r[best]) > 0.998 * sum(r['base']) / len(r['base']):
    best = 'base'
print(best)
PY
)
echo "best: $BEST" | tee gpurun_out/${T}_best.txt
if [ $BEST != base ]; then export CSM_B200_LIB=$L/libcsm_b200_$BEST.so; fi
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
grep -E "passed|failed|rc=|real" gpurun_out/${T}_tests.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_smoke.log
