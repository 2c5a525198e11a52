#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
echo "== tail diag"; timeout 900 python tools/gpu_tail_diag.py 2>&1 | tail -60
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== bench host"; timeout 600 python bench.py --mode host --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_host.json | cut -c1-300
echo "== bake512 host"; timeout 600 python bench.py --workload bake512 --mode host --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02_bench_bake512_host.json | cut -c1-300
