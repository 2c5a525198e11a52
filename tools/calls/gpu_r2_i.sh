#!/bin/bash
# round 2, call I: HOST mode with per-group guarded replay -- parity, benches, light ncu metrics
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu.log
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== bench host"; timeout 600 python bench.py --mode host --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_host.json | cut -c1-300
echo "== bake512 host"; timeout 600 python bench.py --workload bake512 --mode host --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02_bench_bake512_host.json | cut -c1-300
echo "== bench hybrid_host j0"; timeout 600 python bench.py --mode hybrid_host --jitter 0 --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_host_j0.json | cut -c1-300
M=gpu__time_duration.sum,smsp__cycles_active.avg,sm__cycles_elapsed.avg,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers
echo "== ncu metrics: bake host 256^3 + render host"
timeout 900 ncu --metrics $M --clock-control none -k regex:'bake_kernel|render_kernel' -c 3 --csv --log-file gpurun_out/r02_host_metrics.csv python bench.py --mode host --steps 1 --warmup 1 $B > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.DictReader(l for l in open("gpurun_out/r02_host_metrics.csv") if l.startswith('"')))
for r in rows:
    if r["ID"] == rows[-1]["ID"]: print("%-80s %s %s" % (r["Metric Name"], r["Metric Value"], r["Metric Unit"]))
PY
