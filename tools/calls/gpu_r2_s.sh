#!/bin/bash
# round 2, call S: two-ray kernels at 4 blocks per SM (128 registers) for periods <= 8 (Q) against 3 blocks (N2)
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_of_theoretical'))"; }
cp lyapunov3d_b200/liblyap_b200.so /tmp/keep.so
for v in Q N2 Q N2; do
  cp variants/lib$v.so lyapunov3d_b200/liblyap_b200.so
  echo "== $v fast frame"; timeout 600 python bench.py --mode fast $B 2>&1 | tail -1 | pick
  echo "== $v hybrid j0"; timeout 600 python bench.py --mode hybrid --jitter 0 $B 2>&1 | tail -1 | pick
  echo "== $v hybrid_host j0"; timeout 600 python bench.py --mode hybrid_host --jitter 0 $B 2>&1 | tail -1 | pick
done
for v in Q N2; do
  cp variants/lib$v.so lyapunov3d_b200/liblyap_b200.so
  echo "== $v tail diag"; timeout 600 python tools/gpu_tail_diag2.py 2>&1 | grep '"tail_compaction": 1' | grep -v '"exact"'
done
cp /tmp/keep.so lyapunov3d_b200/liblyap_b200.so
