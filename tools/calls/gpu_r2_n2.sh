#!/bin/bash
# round 2, two GPUs: the driver's launch line for bench.py at N=2 and the multi-GPU test
mkdir -p gpurun_out/final
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/final/bench_exact_n2.json; cut -c1-400 gpurun_out/final/bench_exact_n2.json
echo "== pytest multi-gpu"; timeout 900 python -m pytest tests -m gpu -q -x -k "multi_gpu or python_cli" 2>&1 | tail -4 | tee gpurun_out/final/pytest_multigpu.log
