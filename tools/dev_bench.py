"""Developer timing loop (not the contract bench): device-resident sort of n pairs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import inplacemsdradixsort_b200 as m
from bench import parse_count

n = parse_count(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
param = parse_count(sys.argv[3]) if len(sys.argv) > 3 else 0
sched = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else None
if sched:
    m.set_schedule(sched)
src_k, src_r = m.DeviceArray(n), m.DeviceArray(n)
dk, dr = m.DeviceArray(n), m.DeviceArray(n)
m.fill(src_k, src_r, kind=kind, seed=1, param=param)
_, s0, d0 = src_k.check(src_r)
best = None
for it in range(4):
    dk.copy_from(src_k)
    dr.copy_from(src_r)
    ph = m.sort_device(dk.ptr, dr.ptr, n, timed=True)
    tot = sum(ph.values())
    best = tot if best is None else min(best, tot)
    print(it, ph, "total_us", tot, "Gpairs/s %.2f" % (n / tot / 1e3))
bad, s1, d1 = dk.check(dr)
print("n", n, "kind", kind, "sched", m.get_schedule(n), "bad", bad, "sum_ok", s0 == s1, "digest_ok", d0 == d1)
print("levels(us)", [(l["histogram"], l["plan"], l["scatter"]) for l in m.last_level_times()][:6])
print("stats", m.last_stats())
print("best Gpairs/s %.2f" % (n / best / 1e3))
