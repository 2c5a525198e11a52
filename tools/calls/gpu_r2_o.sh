#!/bin/bash
# A/B: queue-head peek (C) against the plain volatile shared flag (B), same box, interleaved
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
cp lyapunov3d_b200/liblyap_b200.so /tmp/keep.so
for rep in 1 2; do for v in C B; do
  cp variants/lib$v.so lyapunov3d_b200/liblyap_b200.so
  echo "== $v exact"; timeout 600 python bench.py $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['peak'], d['roofline']['frac'])"
done; done
for v in C B; do
  cp variants/lib$v.so lyapunov3d_b200/liblyap_b200.so
  echo "== $v fast"; timeout 600 python bench.py --mode fast $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
  echo "== $v hybrid"; timeout 600 python bench.py --mode hybrid --jitter 0 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
  echo "== $v tail diag"; timeout 600 python tools/gpu_tail_diag2.py 2>&1 | grep '"tail_compaction": 1' | grep '"world": 8'
done
cp /tmp/keep.so lyapunov3d_b200/liblyap_b200.so
echo "== pytest (C)"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
