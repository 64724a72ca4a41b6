"""Generate tests/golden/msb64_golden.npz from the UNMODIFIED reference (oracle/_ref).

Run where /root/reference exists (the build container):
    python tests/golden/make_golden.py
The fixtures pin the oracle (oracle/msb64_oracle.c) to outputs of the reference's own
functions; tests/test_oracle_pin.py replays them on any box, with or without the
reference library.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefLib  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "msb64_golden.npz")


def main():
    ref = RefLib()
    g = {}
    # rand.c generator: first values for a few seeds (rand64_init / rand64_next)
    for seed in (0, 1, 42, 0xDEADBEEF):
        g[f"rand64_seed{seed}"] = ref.rand64(seed, 700)          # crosses the 312 refill twice
    # mulhi, binary_search_64
    rng = np.random.default_rng(1)
    x = rng.integers(0, 1 << 64, size=200, dtype=np.uint64)
    y = rng.integers(0, 1 << 64, size=200, dtype=np.uint64)
    g["mulhi_x"], g["mulhi_y"] = x, y
    g["mulhi_out"] = np.array([ref.lib.mulhi(int(a), int(b)) for a, b in zip(x, y)], dtype=np.uint64)
    delim = np.sort(rng.integers(0, 1 << 64, size=128, dtype=np.uint64))
    probes = np.concatenate([delim[::7], rng.integers(0, 1 << 64, size=100, dtype=np.uint64),
                             np.array([0, 0xFFFFFFFFFFFFFFFF], dtype=np.uint64)])
    g["bs_delim"], g["bs_probe"] = delim, probes
    import ctypes as C
    u64p = C.POINTER(C.c_uint64)
    g["bs_out"] = np.array([ref.lib.binary_search_64(delim.ctypes.data_as(u64p), 128, int(p))
                            for p in probes], dtype=np.uint64)
    # schedule_passes over the whole size range the reference accepts
    sizes = [1, 20, 21, 100, 6500, 6501, 10000, 52000, 208000, 208001, 1 << 20, 3_000_000,
             1 << 22, 1 << 23, 1 << 24, 26_000_000, 1 << 25, 1 << 26, 1 << 27, 1 << 28, 1 << 30]
    sched = []
    for sz in sizes:
        p, rb, bf = ref.schedule_passes(sz, 58)
        row = [sz, p] + rb + [0] * (8 - len(rb)) + bf + [0] * (8 - len(bf))
        sched.append(row)
    g["schedule"] = np.array(sched, dtype=np.int64)
    # leaves: insertsort / combsort on small arrays with duplicates
    for name, fn, n in (("insertsort", ref.lib.insertsort, 20), ("combsort", ref.lib.combsort, 300)):
        k = rng.integers(0, 50, size=n, dtype=np.uint64) << np.uint64(40)
        r = np.arange(n, dtype=np.uint64)
        g[f"{name}_in_k"], g[f"{name}_in_r"] = k.copy(), r.copy()
        fn(k.ctypes.data_as(u64p), r.ctypes.data_as(u64p), n)
        g[f"{name}_out_k"], g[f"{name}_out_r"] = k, r
    # histogram + unbuffered in-place partition (deterministic permutation)
    n = 5000
    k = ref.aligned(n)
    k[:] = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
    r = ref.aligned(n)
    r[:] = np.arange(n, dtype=np.uint64)
    g["part_in_k"] = k.copy()
    for shift, bits in ((58, 6), (20, 8), (0, 4)):
        cnt = ref.aligned(1 << bits)
        ref.lib.histogram(k.ctypes.data_as(u64p), n, cnt.ctypes.data_as(u64p), shift, bits)
        g[f"hist_{shift}_{bits}"] = cnt.copy()
    kk, rr = ref.aligned(n), ref.aligned(n)
    kk[:], rr[:] = k, r
    cnt = ref.aligned(64)
    offs = ref.aligned(64)
    ref.lib.histogram(kk.ctypes.data_as(u64p), n, cnt.ctypes.data_as(u64p), 58, 6)
    ref.lib.partition_ip(kk.ctypes.data_as(u64p), rr.ctypes.data_as(u64p), n,
                         cnt.ctypes.data_as(u64p), offs.ctypes.data_as(u64p), 58, 6)
    g["part_out_k"], g["part_out_r"] = kk.copy(), rr.copy()
    # extract_delimiters on a sorted sample with repetitions
    sample = np.sort(np.concatenate([rng.integers(0, 1 << 64, size=900, dtype=np.uint64),
                                     np.full(100, 12345, dtype=np.uint64)]))
    d = np.zeros(64, dtype=np.uint64)
    d[63] = np.uint64(0xFFFFFFFFFFFFFFFF)
    ref.lib.extract_delimiters(sample.ctypes.data_as(u64p), sample.size, d.ctypes.data_as(u64p))
    g["delim_sample"], g["delim_out"] = sample, d
    # range_histogram: range index of every key against 128 delimiters
    rd = ref.aligned(128)
    rd[:] = np.sort(np.concatenate([d[:63], np.array([(p << 58) - 1 for p in range(1, 64)], dtype=np.uint64),
                                    np.array([0xFFFFFFFFFFFFFFFF] * 2, dtype=np.uint64)]))
    keys = ref.aligned(4096)
    keys[:] = rng.integers(0, 1 << 64, size=4096, dtype=np.uint64)
    ranges = np.zeros(4096, dtype=np.uint8)
    count = np.zeros(128, dtype=np.uint64)
    ref.lib.range_histogram(keys.ctypes.data_as(u64p), ranges.ctypes.data_as(C.POINTER(C.c_uint8)),
                            4096, count.ctypes.data_as(u64p), rd.ctypes.data_as(u64p))
    g["rh_delim"], g["rh_keys"], g["rh_ranges"], g["rh_count"] = rd.copy(), keys.copy(), ranges, count
    # local_radixsort (the recursive descent) on a range of 58 significant bits
    for n in (15, 5000, 60000):
        k, r = ref.aligned(n), ref.aligned(n)
        k[:] = rng.integers(0, 1 << 58, size=n, dtype=np.uint64)
        k[: n // 5] = k[0]                                   # duplicates
        r[:] = np.arange(n, dtype=np.uint64)
        g[f"lrs{n}_in_k"] = k.copy()
        ref.local_sort_range(k, r, 58)
        g[f"lrs{n}_out_k"], g[f"lrs{n}_out_r"] = k.copy(), r.copy()
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
