#!/bin/bash
# round 2, call J (2 GPUs): the multi-GPU test and bench.py under torchrun exactly as the driver launches it
mkdir -p gpurun_out
echo "== multi-GPU pytest"; timeout 900 python -m pytest tests -m gpu -q -x -k "multi_gpu or dist or shard" 2>&1 | tail -5 | tee gpurun_out/r02_pytest_multigpu.log
echo "== bench N=2 (driver launch line)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/r02_bench_exact_n2.json | cut -c1-1500
echo "== bench reference arm under torchrun N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -2 | cut -c1-400
