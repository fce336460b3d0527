#!/bin/bash
# key ranges of the flash-decoding attention: 8 (default build) vs 16 (variant build), config 3 decode
T=${1:-r2as}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_as8.json 2> gpurun_out/${T}_bench8.err
CSM_B200_LIB=$PWD/tools/variants/libcsm_as16.so timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_as16.json 2> gpurun_out/${T}_bench16.err
