#!/bin/bash
# round 2, call G: HOST-mode logf variants (tools/probe_hostlog.cu) timed and profiled
mkdir -p gpurun_out build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o build/probe_hostlog tools/probe_hostlog.cu || exit 1
echo "== probe"; ./build/probe_hostlog 1000 2>&1 | tee gpurun_out/r02_probe_hostlog.log
echo "== ncu probe (one launch per variant)"
M=smsp__cycles_active.avg,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_active,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct
timeout 900 ncu --metrics $M --clock-control none --launch-skip 0 --csv --log-file gpurun_out/r02_probe_hostlog_ncu.csv ./build/probe_hostlog 100 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.DictReader(l for l in open("gpurun_out/r02_probe_hostlog_ncu.csv") if l.startswith('"')))
per = collections.OrderedDict()
for r in rows:
    per.setdefault((r["ID"], r["Kernel Name"]), {})[r["Metric Name"]] = r["Metric Value"]
seen = set()
for (i, k), m in per.items():
    if k in seen: continue
    seen.add(k)
    print(k)
    for a, b in m.items(): print("   %-80s %s" % (a, b))
PY
