#!/bin/bash
# skinny GEMM: cluster K split and the NT=4 register cap, A/B on one box
T=${1:-r2ks}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "serving or frame or fullsize or generator or linear or skinny or stress" > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log
echo "== default (two CTAs / SM at NT=4, K split on)" >> gpurun_out/${T}_decode.txt
PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 2 8 16 32 >> gpurun_out/${T}_decode.txt 2>&1
echo "== CSM_SK_KSPLIT=0" >> gpurun_out/${T}_decode.txt
CSM_SK_KSPLIT=0 PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 2 8 16 32 >> gpurun_out/${T}_decode.txt 2>&1
echo "== one CTA / SM at NT=4 (variant build), K split on" >> gpurun_out/${T}_decode.txt
CSM_B200_LIB=$PWD/tools/variants/libcsm_minb1.so PF_SHORT=1 timeout 600 python tools/bench_decode_batch.py 24 32 >> gpurun_out/${T}_decode.txt 2>&1
cat gpurun_out/${T}_decode.txt
