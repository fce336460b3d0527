#!/bin/bash
T=${1:-r2ts2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "mimi" > gpurun_out/${T}_tests.log 2>&1
tail -5 gpurun_out/${T}_tests.log
