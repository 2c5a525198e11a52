#!/bin/bash
# round 2, call R: cheaper per-evaluation prologue/epilogue of the packed fast evaluator (N2) against the previous build (N)
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_of_theoretical'))"; }
cp lyapunov3d_b200/liblyap_b200.so /tmp/keep.so
for v in N2 N N2 N; do
  cp variants/lib$v.so lyapunov3d_b200/liblyap_b200.so
  echo "== $v fast frame"; timeout 600 python bench.py --mode fast $B 2>&1 | tail -1 | pick
  echo "== $v hybrid j0"; timeout 600 python bench.py --mode hybrid --jitter 0 $B 2>&1 | tail -1 | pick
  echo "== $v bake512 fast"; timeout 600 python bench.py --workload bake512 --mode fast $B 2>&1 | tail -1 | pick
done
cp /tmp/keep.so lyapunov3d_b200/liblyap_b200.so
echo "== pytest (N2)"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
