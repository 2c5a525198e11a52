#!/bin/bash
# round 2, final single-GPU measurements on the final build: tests, smoke, every bench line kept under profiles/
mkdir -p gpurun_out/final
O=gpurun_out/final
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench exact (headline)"; timeout 900 python bench.py 2>&1 | tail -1 > $O/bench_exact.json; cut -c1-300 $O/bench_exact.json
for m in fast host; do echo "== bench $m"; timeout 600 python bench.py --mode $m $B $([ $m = host ] && echo --steps 2) 2>&1 | tail -1 > $O/bench_$m.json; cut -c1-200 $O/bench_$m.json; done
for m in hybrid exact hybrid_host; do echo "== bench $m j0"; timeout 600 python bench.py --mode $m --jitter 0 $B 2>&1 | tail -1 > $O/bench_${m}_j0.json; cut -c1-200 $O/bench_${m}_j0.json; done
for m in fast exact host; do echo "== bake512 $m"; timeout 600 python bench.py --workload bake512 --mode $m $B 2>&1 | tail -1 > $O/bench_bake512_$m.json; cut -c1-200 $O/bench_bake512_$m.json; done
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > $O/bench_reference.json; cut -c1-200 $O/bench_reference.json
echo "== longseq"; timeout 900 python tools/gpu_longseq.py 2>&1 | tee $O/longseq.log | cut -c1-220
echo "== tail diag"; timeout 600 python tools/gpu_tail_diag2.py 2>&1 | tee $O/tail_diag2.log | grep '"tail_compaction": 1'
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_exact.csv python bench.py --steps 2 --warmup 1 $B > $O/ncu_launches.log 2>&1
tail -3 $O/launches_bench_exact.csv | cut -c1-200
