"""Developer timing of one sub-range sort as the sharded path runs it: n pairs whose keys lie in
the low `width` bits, sorted with msb64_b200_sort_device_range (phases timed), and the same
sort enqueued `reps` times back to back untimed (what a rank does with its sub-ranges)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import inplacemsdradixsort_b200 as m
from bench import parse_count

n = parse_count(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
width = int(sys.argv[2]) if len(sys.argv) > 2 else 57
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 16
lo, hi = 0, (1 << width) - 1
src_k, src_r = m.DeviceArray(n), m.DeviceArray(n)
m.fill(src_k, src_r, kind=1, seed=1, param=hi)
_, s0, d0 = src_k.check(src_r)
print("n", n, "width", width, "schedule", m.get_range_schedule(n, lo, hi))
ks = [m.DeviceArray(n) for _ in range(reps)]
rs = [m.DeviceArray(n) for _ in range(reps)]
for it in range(3):
    ks[0].copy_from(src_k)
    rs[0].copy_from(src_r)
    ph = m.sort_device(ks[0].ptr, rs[0].ptr, n, timed=True, key_range=(lo, hi))
    print(it, ph, "total_us", sum(ph.values()), "levels", [(l["histogram"], l["plan"], l["scatter"]) for l in m.last_level_times()][:4])
bad, s1, d1 = ks[0].check(rs[0])
print("bad", bad, "sum_ok", s0 == s1, "digest_ok", d0 == d1)
lib = m.load_library()
for it in range(3):
    for k, r in zip(ks, rs):
        k.copy_from(src_k)
        r.copy_from(src_r)
    lib.msb64_b200_stream_sync(None)
    t0 = time.perf_counter()
    for k, r in zip(ks, rs):
        m.sort_device(k.ptr, r.ptr, n, key_range=(lo, hi))
    lib.msb64_b200_stream_sync(None)
    dt = time.perf_counter() - t0
    print(f"{reps} sorts back to back: {dt * 1e3:.2f} ms, {dt * 1e3 / reps:.3f} ms each, {reps * n / dt / 1e9:.2f} Gpairs/s")
