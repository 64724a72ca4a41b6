#!/usr/bin/env python
"""Per-kernel instruction counts from the SASS of the in-tree library.

    python tools/sass_summary.py [ROUND]        -> profiles/rNN_sass_summary.txt

Counts the mnemonics that show which hardware path a kernel takes: UBLKCP (cp.async.bulk, the
non-tensor TMA path), SYNCS (mbarrier), ATOMS (shared atomics), ATOMG / RED (global atomics),
LDG / STG by width, LDS / STS, BAR (block barriers).  Needs only cuobjdump (no GPU).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "inplacemsdradixsort_b200", "lib", "libmsb64_b200.so")
COLS = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "RED", "LDG.64", "LDG.128", "STG.64", "STG.128",
        "LDS", "STS", "BAR", "total"]


def classify(op):
    if op.startswith("UBLKCP"):
        return "UBLKCP"
    if op.startswith("SYNCS"):
        return "SYNCS"
    if op.startswith("ATOMS"):
        return "ATOMS"
    if op.startswith("ATOMG") or op.startswith("ATOM."):
        return "ATOMG"
    if op.startswith("RED"):
        return "RED"
    if op.startswith("LDG"):
        return "LDG.128" if ".128" in op else "LDG.64" if ".64" in op else None
    if op.startswith("STG"):
        return "STG.128" if ".128" in op else "STG.64" if ".64" in op else None
    if op.startswith("LDS"):
        return "LDS"
    if op.startswith("STS"):
        return "STS"
    if op.startswith("BAR"):
        return "BAR"
    return None


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                           capture_output=True, text=True).stdout.splitlines()
    kernels, cur, arch = collections.OrderedDict(), None, set()
    it = iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = next(it)
            name = name.replace("(anonymous namespace)::", "")
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("msb64::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur["total"] += 1
            c = classify(m.group(1))
            if c:
                cur[c] += 1
    out = os.path.join(ROOT, "profiles", f"r{rnd}_sass_summary.txt")
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (arch: {', '.join(sorted(arch))}); static "
                f"instruction counts per kernel\n")
        f.write("# UBLKCP = cp.async.bulk (non-tensor TMA), SYNCS = mbarrier ops, ATOMS = shared atomics, "
                "ATOMG/RED = global atomics\n")
        f.write(f"{'kernel':40s}" + "".join(f"{c:>9s}" for c in COLS) + "\n")
        for k, cnt in kernels.items():
            f.write(f"{k:40s}" + "".join(f"{cnt.get(c, 0):9d}" for c in COLS) + "\n")
    print(out)


if __name__ == "__main__":
    main()
