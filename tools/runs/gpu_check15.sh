#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r2q}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${T}_gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 --fast > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
echo "rc=$?"; tail -c 1500 gpurun_out/${T}_bench_2gpu.json; tail -5 gpurun_out/${T}_bench_2gpu.err
python -m pytest tests/test_parallel_gloo.py tests/test_serving.py -q 2>&1 | tail -2
