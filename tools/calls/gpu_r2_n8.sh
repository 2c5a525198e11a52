#!/bin/bash
# round 2, eight GPUs: the driver's launch line for bench.py at N=8 (strong scaling, shard checks inside the run)
mkdir -p gpurun_out/final
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/final/bench_exact_n8.json; cut -c1-300 gpurun_out/final/bench_exact_n8.json
