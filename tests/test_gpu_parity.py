"""Parity of the CUDA path (through the C ABI) against the CPU oracle, bit-exact.

Oracle = oracle/msb64_oracle.c (restated msb_64.c).  Keys must match the oracle's
sorted key sequence exactly; rids must carry the same (key, rid) multiset (MSD radix
sort is not stable, neither here nor in the reference), which is checked exactly at
these sizes: sorting rids inside every run of equal keys on both sides and comparing.
"""
import os
import sys

import numpy as np
import pytest

from inputs import KINDS, make

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 3, 17, 255, 256, 257, 4095, 4096, 4097, 8191, 8193, 12289, 65536 + 3,
         300_001, 1 << 20]


def canon(keys, rids):
    """Order rids inside equal-key runs so that two correct outputs compare equal."""
    order = np.lexsort((rids, keys))
    return keys[order], rids[order]


def oracle_sorted(oracle, keys, rids):
    k, r = keys.copy(), rids.copy()
    if k.size:
        # whole sort of the reference: sample, 128 ranges, local MSB radix sort per range
        pad_k = np.concatenate([k, np.zeros(k.size // 2 + 64, dtype=np.uint64)])
        pad_r = np.concatenate([r, np.zeros(k.size // 2 + 64, dtype=np.uint64)])
        oracle.sort([pad_k], [pad_r], [k.size])
        k, r = pad_k[:k.size].copy(), pad_r[:k.size].copy()
    return k, r


def run_case(gpu, oracle, keys, rids):
    n = keys.size
    gk, gr = keys.copy(), rids.copy()
    gpu.sort_pairs(gk, gr)
    ok, orr = oracle_sorted(oracle, keys, rids)
    assert np.array_equal(gk, ok), "keys differ from the oracle's sorted keys"
    ck, cr = canon(gk, gr)
    ek, er = canon(ok, orr)
    assert np.array_equal(cr, er), "(key, rid) multiset differs from the oracle's"
    assert n == 0 or np.all(gk[:-1] <= gk[1:])


@pytest.mark.parametrize("n", SIZES)
def test_uniform_sizes(gpu, oracle, n):
    keys = make("uniform", n, seed=n + 1)
    run_case(gpu, oracle, keys, np.arange(n, dtype=np.uint64))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", [5000, 100_003, 1 << 19])
def test_distributions(gpu, oracle, kind, n):
    keys = make(kind, n, seed=7)
    rids = np.arange(n, dtype=np.uint64) * np.uint64(3) + np.uint64(11)
    run_case(gpu, oracle, keys, rids)


@pytest.mark.parametrize("sched", [[8] * 8, [4] * 16, [11, 11, 11, 11, 11, 9], [10, 10, 8, 8, 8, 8, 8, 4],
                                   [9, 9, 9, 9, 9, 9, 10], [5, 6, 7, 8, 9, 10, 11, 8]])
@pytest.mark.parametrize("kind", ["uniform", "dup1000", "low24", "skew", "clustered"])
def test_schedules(gpu, oracle, sched, kind):
    n = 250_007
    keys = make(kind, n, seed=3)
    gpu.set_schedule(sched)
    try:
        assert gpu.get_schedule(n) == sched
        run_case(gpu, oracle, keys, np.arange(n, dtype=np.uint64))
    finally:
        gpu.set_schedule(None)


def test_keys_equal_rids_reference_check(gpu):
    """The reference benchmark's own acceptance test: rids = keys, then check(..., same=1)
    (msb_64.c:2470): ascending keys, key == rid everywhere, checksum = sum of keys."""
    n = 1 << 21
    keys = make("uniform", n, seed=99)
    rids = keys.copy()
    expect = int(np.sum(keys, dtype=np.uint64))
    size = [n]
    gpu.sort([keys], [rids], size)
    assert size == [n]
    assert gpu.check([keys], [rids], size, same=True) == expect


def test_numa_nodes(gpu, oracle):
    """sort() over several arrays: node n gets the n-th key range, equal keys are never
    split, sizes are updated (msb_64.c:2180, 1596-1606)."""
    numa, per = 4, 200_000
    fudge = 1.5
    rng = np.random.default_rng(5)
    keys = [np.concatenate([make("dup1000", per, seed=10 + i), np.zeros(per, dtype=np.uint64)])
            for i in range(numa)]
    rids = [np.concatenate([rng.integers(0, 1 << 64, size=per, dtype=np.uint64),
                            np.zeros(per, dtype=np.uint64)]) for i in range(numa)]
    all_k = np.concatenate([k[:per] for k in keys])
    all_r = np.concatenate([r[:per] for r in rids])
    size = [per] * numa
    gpu.sort(keys, rids, size, numa=numa, fudge=fudge)
    assert sum(size) == numa * per
    out_k = np.concatenate([keys[i][:size[i]] for i in range(numa)])
    out_r = np.concatenate([rids[i][:size[i]] for i in range(numa)])
    assert np.array_equal(out_k, np.sort(all_k))
    ck, cr = canon(out_k, out_r)
    ek, er = canon(all_k, all_r)
    assert np.array_equal(cr, er)
    for i in range(numa - 1):
        if size[i] and size[i + 1]:
            assert keys[i][size[i] - 1] < keys[i + 1][0], "equal keys split across nodes"
    gpu.check(keys, rids, size, numa=numa)


def test_capacity_error(gpu):
    """All-equal keys cannot be spread over two nodes: the reference asserts, we report."""
    per = 10_000
    keys = [np.zeros(per, dtype=np.uint64) for _ in range(2)]
    rids = [np.zeros(per, dtype=np.uint64) for _ in range(2)]
    with pytest.raises(gpu.Msb64Error) as e:
        gpu.sort(keys, rids, [per, per], numa=2, fudge=1.0)
    assert e.value.code == -4


def test_device_resident_and_digest(gpu, oracle):
    """Device API + device-side check: sortedness, checksum and the order-independent
    (key, rid) digest equal the oracle's on the same input."""
    n = (1 << 22) + 12345
    with gpu.DeviceArray(n) as dk, gpu.DeviceArray(n) as dr:
        gpu.fill(dk, dr, kind=0, seed=42)
        keys, rids = dk.download(), dr.download()
        before = oracle.pair_digest(keys, rids)
        phases = gpu.sort_device(dk.ptr, dr.ptr, n, timed=True)
        assert set(phases) == {"histogram", "plan", "scatter", "local_sort", "copy_home", "tail"}
        bad, csum, digest = dk.check(dr)
        assert bad == 0
        assert digest == before
        assert csum == int(np.sum(keys, dtype=np.uint64))
        out = dk.download()
        assert np.array_equal(out, np.sort(keys))
        stats = gpu.last_stats()
        assert stats["error"] == 0


RANGES = [
    (0, (1 << 64) - 1), (5, 5), (7, 8), (0x1234_5678_0000_0000, 0x1234_5678_0000_0FFF),
    (0x7FFF_FFFF_FFFF_FF00, 0x8000_0000_0000_00FF), ((1 << 61), (1 << 62) - 1),
    (3 << 61, (4 << 61) - 1), (0xABC << 52, (0xDEF << 52) + 12345), (1, 1 << 33),
    ((1 << 64) - 1000, (1 << 64) - 1), (0x00FF_0000_0000_0001, 0x0100_0000_0000_0000),
]


@pytest.mark.parametrize("lo,hi", RANGES)
@pytest.mark.parametrize("n", [2, 4097, 300_001, 1 << 21])
def test_sort_with_known_key_range(gpu, lo, hi, n):
    """msb64_b200_sort_device_range: origin-relative first digit, schedule made for the span."""
    rng = np.random.default_rng(n ^ (lo & 0xFFFF))
    span = hi - lo
    off = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
    keys = (np.uint64(lo) + (off % np.uint64(span + 1) if span < (1 << 64) - 1 else off)).astype(np.uint64)
    keys[0], keys[-1] = np.uint64(lo), np.uint64(hi)               # both ends of the range occur
    rids = np.arange(n, dtype=np.uint64) + np.uint64(77)
    with gpu.DeviceArray(n) as dk, gpu.DeviceArray(n) as dr:
        dk.upload(keys)
        dr.upload(rids)
        gpu.sort_device(dk.ptr, dr.ptr, n, key_range=(lo, hi))
        gk, gr = dk.download(), dr.download()
    order = np.lexsort((rids, keys))
    assert np.array_equal(gk, keys[order])
    assert np.array_equal(gr[np.lexsort((gr, gk))], rids[order])


FULL = [("uniform", 0, 0), ("few_distinct_16", 2, 16), ("heavy_duplicates_1e6", 2, 1_000_000),
        ("low_24_bits", 1, 0xFFFFFF), ("presorted", 3, 1), ("reverse_sorted", 4, 1)]


@pytest.mark.parametrize("name,kind,param", FULL, ids=[f[0] for f in FULL])
def test_full_size_properties(gpu, name, kind, param):
    """BASELINE.json's full size (2^30 pairs, configs[1]-[3]) through size-independent
    properties: ascending, wrapping key sum and order-independent (key, rid) digest unchanged
    (the reference's check(), msb_64.c:2432-2505), and idempotence (sorting the sorted
    output changes nothing but possibly the order of rids among equal keys)."""
    n = 1 << 30
    with gpu.DeviceArray(n) as dk, gpu.DeviceArray(n) as dr:
        gpu.fill(dk, dr, kind=kind, seed=2026, param=param)
        bad0, sum0, dig0 = dk.check(dr)
        gpu.sort_device(dk.ptr, dr.ptr, n)
        bad1, sum1, dig1 = dk.check(dr)
        assert bad1 == 0, f"{bad1} descents"
        assert (sum1, dig1) == (sum0, dig0), "key checksum / (key, rid) multiset changed"
        assert gpu.last_stats()["error"] == 0
        gpu.sort_device(dk.ptr, dr.ptr, n)
        assert dk.check(dr) == (0, sum0, dig0), "not idempotent"


def test_random_sizes_and_kinds(gpu):
    """Many ragged sizes x input families (unit tails, odd offsets, arrays ending inside a
    tile): keys must equal the sorted keys and the (key, rid) multiset must be unchanged.
    The expected order comes from numpy here (the oracle's order is pinned to it elsewhere)."""
    rng = np.random.default_rng(20261018)
    for case in range(120):
        kind = KINDS[int(rng.integers(0, len(KINDS)))]
        n = int(rng.choice([rng.integers(1, 300), rng.integers(300, 9000), rng.integers(9000, 400_000)]))
        keys = make(kind, n, seed=1000 + case)
        rids = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
        gk, gr = keys.copy(), rids.copy()
        gpu.sort_pairs(gk, gr)
        order = np.lexsort((rids, keys))
        assert np.array_equal(gk, keys[order]), f"case {case}: {kind} n={n}: keys"
        assert np.array_equal(gr[np.lexsort((gr, gk))], rids[order]), f"case {case}: {kind} n={n}: rids"


def test_device_sort_at_odd_offsets(gpu):
    """Device arrays that start and end at odd element offsets inside a larger allocation
    (16-byte alignment of the bulk copies must not leak into the interface: the C ABI asks
    for 16-byte aligned arrays, so the offset is even, but lengths are arbitrary)."""
    n_total = 1_200_007
    base = make("uniform", n_total, seed=77)
    for off, n in [(0, 1_200_007), (2, 1_200_005), (4098, 777_777), (65538, 4097), (8, 2_001)]:
        keys = base[off: off + n].copy()
        rids = np.arange(n, dtype=np.uint64)
        with gpu.DeviceArray(n_total) as dk, gpu.DeviceArray(n_total) as dr:
            dk.upload(base)
            full_r = np.zeros(n_total, dtype=np.uint64)
            full_r[off: off + n] = rids
            dr.upload(full_r)
            gpu.sort_device(dk.ptr + 8 * off, dr.ptr + 8 * off, n)
            out_k, out_r = dk.download(), dr.download()
        assert np.array_equal(out_k[:off], base[:off]) and np.array_equal(out_k[off + n:], base[off + n:]), \
            "the sort touched elements outside its arrays"
        assert np.array_equal(out_k[off: off + n], np.sort(keys))
        got_r = out_r[off: off + n]
        assert np.array_equal(keys[got_r.astype(np.int64)], out_k[off: off + n]), "rid does not point at its key"


@pytest.mark.parametrize("numa,virtual,kind", [(1, 0, 0), (2, 0, 0), (3, 0, 2), (2, 1, 0), (3, 1, 1), (4, 1, 2)])
def test_dropin_c_program(gpu, numa, virtual, kind):
    """A C program that includes the reference's include/msb_64.h verbatim, linked against
    libmsb64_b200.so (tests/c/dropin.c): calls mamalloc() + sort() and checks the result.
    virtual = 1: the nodes are sorted as shards sharing the visible GPUs (the multi-GPU path of
    sort(), msb_64.c:2261-2275 across NUMA nodes) even on a box with one GPU."""
    import subprocess
    sys.path.insert(0, os.path.join(ROOT, "tests", "c"))
    import build_dropin
    path = build_dropin.build()
    assert path and os.path.exists(path), "tests/c/_build/dropin must be built where the reference header is"
    env = dict(os.environ)
    env.pop("MSB64_B200_VIRTUAL_SHARDS", None)
    env["MSB64_B200_SINGLE_DEVICE"] = "1"
    if virtual:
        env["MSB64_B200_VIRTUAL_SHARDS"] = "1"
        env.pop("MSB64_B200_SINGLE_DEVICE")
    res = subprocess.run([path, str(numa), "400003", "1.5", str(kind)], capture_output=True, text=True, env=env,
                         timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "dropin ok" in res.stdout


@pytest.mark.parametrize("kind", ["uniform", "low24", "dup1000", "sorted"])
@pytest.mark.parametrize("numa", [2, 5])
def test_sort_over_shards(gpu, oracle, kind, numa):
    """sort() with numa arrays on numa (virtual) GPU shards: node n gets the n-th key range."""
    per = 150_000
    fudge = 1.6
    keys = [np.concatenate([make(kind, per + 31 * i, seed=70 + i), np.zeros(per, dtype=np.uint64)]) for i in range(numa)]
    rids = [np.concatenate([np.arange(per + 31 * i, dtype=np.uint64) + np.uint64(i << 40), np.zeros(per, dtype=np.uint64)])
            for i in range(numa)]
    size = [per + 31 * i for i in range(numa)]
    all_k = np.concatenate([k[:s] for k, s in zip(keys, size)])
    all_r = np.concatenate([r[:s] for r, s in zip(rids, size)])
    os.environ["MSB64_B200_VIRTUAL_SHARDS"] = "1"
    try:
        phases = gpu.sort(keys, rids, size, numa=numa, fudge=fudge)
    finally:
        del os.environ["MSB64_B200_VIRTUAL_SHARDS"]
    assert any("exchange" in p for p in phases), f"the sharded path did not run: {phases}"
    assert sum(size) == all_k.size
    out_k = np.concatenate([keys[i][:size[i]] for i in range(numa)])
    out_r = np.concatenate([rids[i][:size[i]] for i in range(numa)])
    n = all_k.size
    wk = np.concatenate([all_k, np.zeros(n // 2 + 64, np.uint64)])
    wr = np.concatenate([all_r, np.zeros(n // 2 + 64, np.uint64)])
    oracle.sort([wk], [wr], [n])
    assert np.array_equal(out_k, wk[:n])
    ck, cr = canon(out_k, out_r)
    ek, er = canon(all_k, all_r)
    assert np.array_equal(cr, er)
    gpu.check(keys, rids, size, numa=numa)
