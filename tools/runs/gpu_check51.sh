#!/bin/bash
# final tree, two ranks over NCCL as the driver launches the scaling run (own arm only, --fast secondary workloads)
T=${1:-r2g2f}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --fast > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
head -c 500 gpurun_out/${T}_bench_2gpu.json; echo; tail -3 gpurun_out/${T}_bench_2gpu.err
