#!/usr/bin/env python
"""Turn the ncu outputs of a gpurun call into the tracked summaries under profiles/.

    python tools/profile_summary.py ROUND LAUNCHES.csv REPORT.ncu-rep [REPORT2.ncu-rep ...]

Writes
  profiles/rNN_launches.txt       per-kernel totals of the launch list (gpu__time_duration)
  profiles/rNN_kernels.csv        one row per profiled launch of the --set full captures
  profiles/scatter_traffic.json   dram bytes per scatter launch (bench.py's roofline.traffic)
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("msb64::", "").replace("<unnamed>::", "").replace("void ", "")
    return name.strip()


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    tot = collections.OrderedDict()
    for r in rows:
        k = short(r[4])
        ns = float(r[14])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += ns
    whole = sum(v[1] for v in tot.values())
    ours = sum(v[1] for k, v in tot.items() if re.match(r"(histogram|scatter|plan|local_sort(_packed)?|copy|init|route|bucket_route|tail|digit_histogram)_kernel", k))
    with open(out, "w") as f:
        f.write(f"# {os.path.basename(path)}: {len(rows)} launches, {whole / 1e6:.3f} ms of kernel time "
                f"(ncu serialises launches and runs them cold; shares, not absolutes, carry over)\n")
        f.write(f"# share of the sort's own kernels: {100 * ours / whole:.1f}% (the rest: fill, check, torch copies)\n")
        f.write(f"{'kernel':48s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'share of sort':>14s}\n")
        for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            mine = re.match(r"(histogram|scatter|plan|local_sort(_packed)?|copy|init|route|bucket_route|tail|digit_histogram)_kernel", k)
            f.write(f"{k:48s} {n:8d} {ns / 1e6:10.3f} {100 * ns / whole:6.1f}% "
                    f"{(100 * ns / ours if mine else 0):13.1f}%\n")
    return out


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rnd, launch_csv, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    print(launches(launch_csv, os.path.join(ROOT, "profiles", f"r{rnd}_launches.txt")))
    out_rows, scatter = [], []
    for rep in reps:
        h, units, rows = raw(rep)
        idx = {m: h.index(m) for m in METRICS if m in h}
        for r in rows:
            name = short(r[h.index("Kernel Name")])
            rec = {"report": os.path.basename(rep), "kernel": name, "grid": r[h.index("Grid Size")],
                   "block": r[h.index("Block Size")]}
            for m, i in idx.items():
                rec[m + (f" [{units[i]}]" if units[i] else "")] = r[i]
            out_rows.append(rec)
            if name.startswith("scatter_kernel"):
                u = units[idx["dram__bytes_read.sum"]]
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                b = (float(r[idx["dram__bytes_read.sum"]]) + float(r[idx["dram__bytes_write.sum"]])) * scale
                if b > 1e9:
                    scatter.append((b, float(r[idx["gpu__time_duration.sum"]])))
    keys = list(out_rows[0].keys()) if out_rows else []
    path = os.path.join(ROOT, "profiles", f"r{rnd}_kernels.csv")
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for rec in out_rows:
            w.writerow({k: rec.get(k, "") for k in keys})
    print(path)
    if scatter:
        path = os.path.join(ROOT, "profiles", "scatter_traffic.json")
        json.dump({"dram_bytes_per_launch": sum(b for b, _ in scatter) / len(scatter),
                   "launches": len(scatter), "source": f"ncu --set full, round {rnd}, "
                   "dram__bytes_read.sum + dram__bytes_write.sum of the non-empty scatter launches",
                   "pairs": int(os.environ.get("MSB64_PROFILE_PAIRS", 1 << 30)), "workload": "uniform"},
                  open(path, "w"), indent=1)
        print(path)


if __name__ == "__main__":
    main()
