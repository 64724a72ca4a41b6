#!/usr/bin/env python
"""Per-source-line summary of one kernel of an ncu report (needs -lineinfo at compile time).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_NAME [min_percent] [launch_index]

Prints, for every CUDA source line that holds at least `min_percent` of the warp-stall
samples: share of samples, share of executed instructions, shared-memory excess wavefronts,
and the three most frequent stall reasons.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
           "--kernel-name", kernel]
    if len(sys.argv) > 4:
        cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    lines = []          # (file, line, source, cols...)
    hdr, cur_file = None, ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            lines.append((cur_file, r))
    if not hdr:
        sys.exit("no source view in report (kernel name wrong or no -lineinfo?)")
    ci = {}
    for i, h in enumerate(hdr):
        ci.setdefault(h, i)
    S, I = ci["# Samples"], ci["Instructions Executed"]
    X = ci.get("L1 Wavefronts Shared Excessive")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(r, i):
        try:
            return int(r[i])
        except (ValueError, IndexError, TypeError):
            return 0
    tot = sum(num(r, S) for _, r in lines) or 1
    toti = sum(num(r, I) for _, r in lines) or 1
    print(f"kernel {kernel}: {tot} samples, {toti} warp instructions, {len(lines)} source lines")
    for f, r in lines:
        if num(r, S) < tot * minpct / 100:
            continue
        st = sorted(((num(r, ci[s]), s[6:]) for s in stalls), reverse=True)[:3]
        sts = " ".join(f"{s}:{100 * c // max(num(r, S), 1)}%" for c, s in st if c)
        print(f"{f}:{r[0]:>4} {100 * num(r, S) / tot:5.1f}%smp {100 * num(r, I) / toti:5.1f}%ins "
              f"xwf={num(r, X) if X else 0:>10}  {r[1].strip()[:64]:64s} | {sts}")


if __name__ == "__main__":
    main()
