#!/bin/bash
# round 2, call U: chord-descending tile order -- invariance tests, then timings against image order
echo "== pytest (order / partition / tail / hybrid tests)"; timeout 1200 python -m pytest tests -m gpu -q -x -k "tile_order or tile_partition or tail_compaction or hybrid or multi_gpu or miss_pixels or host_buffer or headless" 2>&1 | tail -5
echo "== timings"; timeout 900 python tools/gpu_tile_order.py 2>&1 | tee gpurun_out/r2_tile_order.log
