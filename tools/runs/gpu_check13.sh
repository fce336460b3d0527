#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2o}
python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_frame.py tests/test_gpu_fullsize.py tests/test_gpu_attn_prefill.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
grep -E "passed|failed|rc=|Error|error|assert" gpurun_out/${T}_tests.log | tail -6
python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/'+"${T}"+'_bench_full.json').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
s=d['secondary']
for k in ('decode_bs1_sampled','config3_B1','config3_B8','config3_B32','config4_mimi_decode'):
    v=s.get(k,{})
    if 'error' in v: print(k, v); continue
    if 'prefill' in v: print(k, 'prefill ms', v['prefill']['ms'], 'frac', round(v['prefill']['roofline']['frac'],3), '| decode ms', v['decode']['ms_per_step'], 'frac', round(v['decode']['roofline']['frac'],3))
    elif 'ms' in v: print(k, v['ms'], round(v['roofline']['frac'],4), round(v['roofline_hbm']['frac'],4))
    else: print(k, v.get('ms_per_step'), v.get('roofline',{}).get('frac'))
print(json.dumps(s.get('config5_256_requests'), indent=0)[:900])
PY
tail -3 gpurun_out/${T}_bench_full.err
