#!/bin/bash
# eight ranks over NCCL, as the driver launches the scaling run (own arm only)
T=${1:-r2g8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_8gpu.json 2> gpurun_out/${T}_bench_8gpu.err
tail -c 600 gpurun_out/${T}_bench_8gpu.json; tail -3 gpurun_out/${T}_bench_8gpu.err
