"""The C-ABI library without a GPU: it loads, exports every symbol include/msb64_b200.h
declares, keeps the reference's signatures, and refuses to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msb64_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
    skip = {"defined", "if", "sizeof"}
    return sorted({n for n in names if n not in skip and not n.isupper()})


def test_header_and_library_agree(msb):
    lib = msb.load_library()
    declared = declared_functions()
    assert "sort" in declared and "mamalloc" in declared
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, f"declared but not exported: {missing}"
    assert sorted(msb.EXPORTS) == declared, "msb64.EXPORTS is out of date with the header"


def test_reference_signatures_are_kept():
    """sort() and mamalloc() read exactly like reference include/msb_64.h:36-40."""
    text = re.sub(r"\s+", " ", open(HEADER).read())
    assert ("void sort(uint64_t **keys, uint64_t **rids, uint64_t *size, int threads, int numa, "
            "double fudge, char **description, uint64_t *times);") in text
    assert "void *mamalloc(size_t size);" in text


def test_library_is_sm100a_only(msb):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", msb.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_mamalloc_alignment(msb):
    a = msb.mamalloc(1000)
    assert a.ctypes.data % 64 == 0 and a.size == 1000
    a[:] = 7
    msb.mafree(a)


def test_no_cpu_fallback(msb):
    """Without a device every compute entry point reports MSB64_ERR_CUDA; nothing sorts."""
    if msb.device_count() > 0:
        pytest.skip("a GPU is visible; the refusal path is for CPU-only boxes")
    k = np.array([3, 1, 2], dtype=np.uint64)
    r = np.arange(3, dtype=np.uint64)
    with pytest.raises(msb.Msb64Error) as e:
        msb.sort_pairs(k, r)
    assert e.value.code == -1
    assert k.tolist() == [3, 1, 2], "input must be untouched"
    with pytest.raises(msb.Msb64Error):
        msb.sort([k], [r], [3])


def test_product_does_not_touch_the_oracle():
    """The package and the CUDA sources never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "inplacemsdradixsort_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in text and "from oracle" not in text \
                    and "import oracle" not in text and "msb64_oracle" not in text, f
    assert "oracle" not in open(HEADER).read().lower()


def test_schedule_api(msb):
    for e in range(0, 33):
        s = msb.get_schedule(1 << e)
        # the digits cover all 64 bits; the last one starts above bit 0 (it is clamped there)
        assert sum(s) >= 64 > sum(s[:-1]) and all(4 <= b <= 11 for b in s) and len(s) <= 16, (e, s)
    msb.set_schedule([8] * 8)
    assert msb.get_schedule(12345) == [8] * 8
    msb.set_schedule(None)
    for bad in ([8] * 7, [3] + [8] * 7 + [5], [12, 12, 12, 12, 12, 4]):
        with pytest.raises(msb.Msb64Error):
            msb.set_schedule(bad)
    assert msb.get_schedule(1 << 30)[:3] == [7, 6, 6]
    # behind the digits uniform keys need: 7-bit tail digits (one cooperative launch, msb64_tail.cuh)
    assert set(msb.get_schedule(1 << 30)[3:]) == {7}
    # the sub-range sorts of the sharded path: 2^25 pairs in a 56-bit range take two 7-bit passes,
    # 2^26 pairs in a 57-bit range 8 + 7 (msb64_shard.cuh)
    assert msb.get_range_schedule(1 << 25, 0, (1 << 56) - 1)[0][:3] == [7, 7, 7]
    assert msb.get_range_schedule(1 << 26, 0, (1 << 57) - 1)[0][:2] == [8, 7]


def test_workspace_bytes(msb):
    prev = 0
    for e in (10, 16, 20, 24, 28, 30):
        w = msb.workspace_bytes(1 << e)
        assert w >= 2 * 8 * (1 << e) and w > prev
        prev = w
    assert msb.workspace_bytes(1 << 30) < 2.2 * 16 * (1 << 30)


def test_range_schedule_invariants(msb):
    """msb64_b200_sort_device_range's host logic (no device): for any key range the first
    digit of the largest key fits its width, widths stay within the kernels' 4..11 bits, and
    the digits cover every bit in which two keys of the range can differ."""
    import random
    rnd = random.Random(7)
    cases = [(0, (1 << 64) - 1), (5, 5), (7, 8), (0, 1), ((1 << 64) - 2, (1 << 64) - 1),
             (0x7FFF_FFFF_FFFF_FFFF, 0x8000_0000_0000_0000), (1 << 63, (1 << 64) - 1)]
    for _ in range(3000):
        w = rnd.randint(0, 64)
        lo = rnd.getrandbits(64)
        span = rnd.getrandbits(w) if w else 0
        cases.append((lo, min(lo + span, (1 << 64) - 1)))
    for lo, hi in cases:
        for n in (2, 5000, 1 << 20, (1 << 30) + 12345, 3 << 30):
            bits, shift0, origin0 = msb.get_range_schedule(n, lo, hi)
            assert bits and all(4 <= b <= 11 for b in bits), (lo, hi, n, bits)
            assert len(bits) <= 16
            assert 0 <= shift0 <= 64 - 4
            assert origin0 == lo >> shift0
            assert (hi >> shift0) - origin0 < (1 << bits[0]), (lo, hi, n, bits, shift0)
            # the digits below the first one are bit fields under shift0 and must reach bit 0
            assert sum(bits[1:]) >= shift0, (lo, hi, n, bits, shift0)
    # the full range reproduces the default schedule
    assert msb.get_range_schedule(1 << 30, 0, (1 << 64) - 1)[0] == msb.get_schedule(1 << 30)


def test_dropin_program_links(msb):
    """tests/c/dropin.c includes the REFERENCE's include/msb_64.h verbatim and links against
    libmsb64_b200.so: sort() and mamalloc() resolve to this library (run: tests/test_gpu_parity.py)."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "c"))
    import build_dropin
    if not os.path.exists(os.path.join(build_dropin.REF_INCLUDE, "msb_64.h")):
        pytest.skip("the reference header is not on this box; the binary was built where it is")
    path = build_dropin.build()
    assert path and os.path.exists(path)
    src = open(os.path.join(ROOT, "tests", "c", "dropin.c")).read()
    assert '#include "msb_64.h"' in src and "msb64_b200" not in src.split("*/", 1)[1], \
        "the program must be written against the reference header only"
    und = subprocess.run(["nm", "-D", "--undefined-only", path], capture_output=True, text=True).stdout
    assert " sort" in und and " mamalloc" in und
    ldd = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "libmsb64_b200.so" in ldd and "not found" not in ldd


def _hist_rows(msb, key_sets, shift=52, bits=12, origin=0):
    slots = msb.load_library().msb64_b200_shard_slots()
    h = np.zeros((len(key_sets), slots), dtype=np.uint64)
    nb = 1 << bits
    for r, k in enumerate(key_sets):
        d = ((k >> np.uint64(shift)) - np.uint64(origin)) & np.uint64(nb - 1)
        h[r, :nb] = np.bincount(d.astype(np.int64), minlength=nb)
        h[r, nb] = k.min() if k.size else np.uint64((1 << 64) - 1)
        h[r, nb + 1] = k.max() if k.size else 0
    return h


@pytest.mark.parametrize("world", [1, 2, 3, 8, 16, 64])
def test_shard_plan_host(msb, world):
    """The plan of the sharded sort (msb64_b200_shard_plan_host; host only): monotone bucket
    table, counts that match the histograms, destinations balanced, sub-ranges balanced."""
    from inplacemsdradixsort_b200 import msb64
    lib = msb.load_library()
    subs = lib.msb64_b200_shard_subs(world)
    want = 32
    while want > 1 and world * want > 128:
        want //= 2
    assert subs == want
    rng = np.random.default_rng(world)
    per = 40_000
    ks = [rng.integers(0, 1 << 64, size=per + 13 * r, dtype=np.uint64) for r in range(world)]
    h = _hist_rows(msb, ks)
    caps = [int(per * 1.3)] * world
    rc, shift, bits, origin, table, counts = msb64.shard_plan_host(h, caps)
    assert (rc, shift, bits, origin) == (0, 52, 12, 0)
    assert table.size == 4096 and np.all(np.diff(table.astype(np.int64)) >= 0) and int(table.max()) < world * subs
    digit = [(k >> np.uint64(52)).astype(np.int64) for k in ks]
    for r in range(world):
        assert np.array_equal(counts[r], np.bincount(table[digit[r]], minlength=world * subs).astype(np.uint64))
    per_dest = counts.sum(axis=0).reshape(world, subs).sum(axis=1)
    total = sum(k.size for k in ks)
    assert per_dest.sum() == total
    assert per_dest.max() - per_dest.min() <= 4 * (total / 4096 + 200)          # bins are never split
    if world <= 8:
        per_sub = counts.sum(axis=0).astype(np.int64)
        assert per_sub.max() - per_sub.min() <= 4 * (total / 4096 + 200)


def test_shard_plan_host_window_and_capacity(msb):
    from inplacemsdradixsort_b200 import msb64
    rng = np.random.default_rng(9)
    world, per = 4, 30_000
    # keys that share their top bits: the first plan asks for a window on the keys' span
    ks = [(rng.integers(0, 1 << 64, size=per, dtype=np.uint64) & np.uint64((1 << 24) - 1)) + np.uint64(1 << 40)
          for _ in range(world)]
    caps = [int(per * 1.3)] * world
    rc, shift, bits, origin, _, _ = msb64.shard_plan_host(_hist_rows(msb, ks), caps)
    assert rc == 1 and bits == 13 and shift == 12 and origin == int(min(k.min() for k in ks)) >> 12
    rc, s2, b2, o2, table, counts = msb64.shard_plan_host(_hist_rows(msb, ks, shift, bits, origin), caps,
                                                          may_retry=False, shift=shift, bits=bits, origin=origin)
    assert (rc, s2, b2, o2) == (0, shift, bits, origin)
    assert table.size == 1 << 13
    per_dest = counts.sum(axis=0).reshape(world, -1).sum(axis=1)
    assert per_dest.max() <= caps[0] and per_dest.sum() == world * per
    # all keys equal: no cut can spread them -- CAPACITY, like the reference's assert (msb_64.c:1574)
    eq = [np.full(per, 77, dtype=np.uint64) for _ in range(world)]
    rc, shift, bits, origin, _, _ = msb64.shard_plan_host(_hist_rows(msb, eq), caps)
    if rc == 1:
        rc = msb64.shard_plan_host(_hist_rows(msb, eq, shift, bits, origin), caps, may_retry=False,
                                   shift=shift, bits=bits, origin=origin)[0]
    assert rc == -4
    assert "capacity" in msb.load_library().msb64_b200_last_error().decode()
    # ranks without keys take part
    some = [ks[0], np.zeros(0, np.uint64), ks[2], np.zeros(0, np.uint64)]
    rc, shift, bits, origin, _, _ = msb64.shard_plan_host(_hist_rows(msb, some), caps)
    assert rc == 1
    rc, _, _, _, _, counts = msb64.shard_plan_host(_hist_rows(msb, some, shift, bits, origin), caps, may_retry=False,
                                                   shift=shift, bits=bits, origin=origin)
    assert rc == 0 and counts[1].sum() == 0 and counts.sum() == 2 * per
