#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small markdown file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_render_exact.ncu-rep profiles/r01_render_exact.md "title"
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = ncu(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}", "", f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, one launch).",
             f"Kernel: `{vals[idx['Kernel Name']]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in idx:
            lines.append(f"| `{k}` | {vals[idx[k]]} | {units[idx[k]]} |")
    src = ncu(rep, "source")
    h = src[1]
    ia, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    by, samp = collections.Counter(), collections.Counter()
    tot = 0
    for r in src[2:]:
        if len(r) <= ia:
            continue
        toks = r[isrc].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        by[op] += int(r[ia])
        samp[op] += int(r[isamp])
        tot += int(r[ia])
    lines += ["", "## Executed warp-instructions by opcode (SASS source page)", "", "| opcode | warp-instructions | share | stall samples |", "|---|---|---|---|"]
    for op, c in by.most_common(14):
        lines.append(f"| {op} | {c} | {100 * c / tot:.2f} % | {samp[op]} |")
    lines.append(f"| total | {tot} | 100 % | {sum(samp.values())} |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
