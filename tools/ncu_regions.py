#!/usr/bin/env python
"""Where the non-FMA instructions of a kernel go: contiguous SASS regions of equal execution count,
ranked by executed warp-instructions (from an .ncu-rep with the source page).

    python tools/ncu_regions.py gpurun_out/x.ncu-rep [min_share_percent] [--list A B]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[1]
    ia, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    ins = [(r[isrc].strip(), int(r[ia]), int(r[isamp])) for r in rows[2:] if len(r) > ia]
    tot = sum(c for _, c, _ in ins)
    if "--list" in sys.argv:
        a, b = int(sys.argv[sys.argv.index("--list") + 1]), int(sys.argv[sys.argv.index("--list") + 2])
        for i in range(a, b):
            print(i, ins[i][1], ins[i][2], ins[i][0])
        return
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
    regs, start = [], 0
    for i in range(1, len(ins) + 1):
        if i == len(ins) or abs(ins[i][1] - ins[start][1]) > 0.02 * max(ins[start][1], 1):
            regs.append((start, i))
            start = i
    print("total warp-instructions", tot)
    for a, b in regs:
        n = sum(c for _, c, _ in ins[a:b])
        packed = sum(c for s, c, _ in ins[a:b] if "FFMA2" in s or "FMUL2" in s)
        if 100.0 * n / tot >= thr:
            print(f"[{a:5d},{b:5d}) len {b - a:4d}  exec/instr {ins[a][1]:>12d}  total {n:>14d} ({100.0 * n / tot:5.2f} %)  packed {packed:>14d}  samples {sum(s for _, _, s in ins[a:b])}")


if __name__ == "__main__":
    main()
