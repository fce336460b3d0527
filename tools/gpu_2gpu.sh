#!/bin/bash
# two ranks over NCCL, as the driver launches the scaling run
T=${1:-r2g2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
tail -c 1500 gpurun_out/${T}_bench_2gpu.json; tail -3 gpurun_out/${T}_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref_2gpu.json 2> gpurun_out/${T}_bench_ref_2gpu.err
tail -c 400 gpurun_out/${T}_bench_ref_2gpu.json
