#!/bin/bash
# round 2, call H: replicated-table HOST mode -- parity, benches, probe, ncu capture
mkdir -p gpurun_out build
echo "== pytest (host-related + all)"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu.log
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== bench host"; timeout 600 python bench.py --mode host --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_host.json | cut -c1-300
echo "== bake512 host"; timeout 600 python bench.py --workload bake512 --mode host --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02_bench_bake512_host.json | cut -c1-300
echo "== bench hybrid_host j0"; timeout 600 python bench.py --mode hybrid_host --jitter 0 --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_host_j0.json | cut -c1-300
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o build/probe_hostlog tools/probe_hostlog.cu && ./build/probe_hostlog 1000 2>&1 | tee gpurun_out/r02_probe_hostlog.log
echo "== ncu full render host"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_render_host python bench.py --mode host --steps 1 --warmup 1 $B > gpurun_out/ncu_render_host.log 2>&1
ls -la gpurun_out | tail -5
