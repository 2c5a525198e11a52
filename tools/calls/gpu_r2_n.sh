#!/bin/bash
# round 2, call N: two-ray kernels at 4 blocks/SM for short periods
mkdir -p gpurun_out
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== bench fast"; timeout 600 python bench.py --mode fast $B 2>&1 | tail -1 | cut -c1-200
echo "== bench hybrid j0"; timeout 600 python bench.py --mode hybrid --jitter 0 $B 2>&1 | tail -1 | cut -c1-200
echo "== bench exact"; timeout 600 python bench.py $B 2>&1 | tail -1 | cut -c1-200
echo "== tail diag"; timeout 600 python tools/gpu_tail_diag2.py 2>&1 | grep '"tail_compaction": 1' | tail -12
