#!/bin/bash
# round 2, last check of the final build: GPU tests, smoke, the default bench line
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/final/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --no-subrecords 2>&1 | tail -1 | cut -c1-260
