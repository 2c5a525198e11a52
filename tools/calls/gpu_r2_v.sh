#!/bin/bash
# round 2, call V: the build with the chord-ordered queue -- all GPU tests, smoke, the bench lines it changes
mkdir -p gpurun_out/final
O=gpurun_out/final
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench exact (headline)"; timeout 900 python bench.py 2>&1 | tail -1 > $O/bench_exact.json; cut -c1-300 $O/bench_exact.json
echo "== bench fast"; timeout 600 python bench.py --mode fast $B 2>&1 | tail -1 > $O/bench_fast.json; cut -c1-200 $O/bench_fast.json
for m in hybrid exact hybrid_host; do echo "== bench $m j0"; timeout 600 python bench.py --mode $m --jitter 0 $B 2>&1 | tail -1 > $O/bench_${m}_j0.json; cut -c1-200 $O/bench_${m}_j0.json; done
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_exact.csv python bench.py --steps 2 --warmup 1 $B > $O/ncu_launches.log 2>&1
grep -c . $O/launches_bench_exact.csv
