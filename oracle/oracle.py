"""ctypes bindings for the CPU oracle (and, when present, the compiled reference).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package.

  Oracle   -> oracle/liboracle_msb64.so   (our restatement, msb64_oracle.c)
  RefLib   -> oracle/_ref/libmsb64_ref.so (unmodified reference sources, built by
              oracle/Makefile where /root/reference exists; prebuilt file otherwise)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle_msb64.so")
REF_SO = os.path.join(HERE, "_ref", "libmsb64_ref.so")

_u64p = C.POINTER(C.c_uint64)
_i8p = C.POINTER(C.c_int8)


def build(force: bool = False) -> None:
    """Compile the oracle (always) and oracle/_ref (only where the reference tree is)."""
    src = os.path.join(HERE, "msb64_oracle.c")
    stale = (not os.path.exists(ORACLE_SO)
             or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "liboracle_msb64.so"])
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _ptr(a: np.ndarray, typ=_u64p):
    return a.ctypes.data_as(typ)


def _check_u64(a: np.ndarray) -> np.ndarray:
    assert a.dtype == np.uint64 and a.flags.c_contiguous, "need contiguous uint64"
    return a


class Oracle:
    """Scalar restatement of msb_64.c (see msb64_oracle.c for the line map)."""

    def __init__(self):
        build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.orc_rand64_fill.argtypes = [C.c_uint64, _u64p, C.c_uint64]
        L.orc_mulhi.restype = C.c_uint64
        L.orc_mulhi.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_binary_search.restype = C.c_uint64
        L.orc_binary_search.argtypes = [_u64p, C.c_uint64, C.c_uint64]
        L.orc_insertsort.argtypes = [_u64p, _u64p, C.c_uint64]
        L.orc_combsort.argtypes = [_u64p, _u64p, C.c_uint64]
        L.orc_histogram.argtypes = [_u64p, C.c_uint64, _u64p, C.c_uint8, C.c_uint8]
        L.orc_partition_ip.argtypes = [_u64p, _u64p, C.c_uint64, _u64p, _u64p,
                                       C.c_uint8, C.c_uint8]
        L.orc_schedule_passes.restype = C.c_int
        L.orc_schedule_passes.argtypes = [C.c_uint64, C.c_int8, _i8p, _i8p]
        L.orc_sort_range.restype = C.c_int
        L.orc_sort_range.argtypes = [_u64p, _u64p, C.c_uint64, C.c_int8]
        L.orc_extract_delimiters.argtypes = [_u64p, C.c_uint64, _u64p]
        L.orc_range_delimiters.argtypes = [_u64p, C.c_uint64, _u64p, _u64p]
        L.orc_range_histogram.argtypes = [_u64p, C.POINTER(C.c_uint8), C.c_uint64,
                                          _u64p, _u64p]
        L.orc_sort.restype = C.c_int
        L.orc_sort.argtypes = [C.POINTER(_u64p), C.POINTER(_u64p), _u64p, _u64p,
                               C.c_int, C.c_uint64]
        L.orc_check_sorted.restype = C.c_uint64
        L.orc_check_sorted.argtypes = [_u64p, C.c_uint64, _u64p]
        L.orc_pair_digest.restype = C.c_uint64
        L.orc_pair_digest.argtypes = [_u64p, _u64p, C.c_uint64]
        L.orc_key_sequence_digest.restype = C.c_uint64
        L.orc_key_sequence_digest.argtypes = [_u64p, C.c_uint64]

    # -- generator ---------------------------------------------------------
    def rand64(self, seed: int, count: int) -> np.ndarray:
        out = np.empty(count, dtype=np.uint64)
        self.lib.orc_rand64_fill(seed, _ptr(out), count)
        return out

    # -- leaves ------------------------------------------------------------
    def insertsort(self, keys, rids):
        self.lib.orc_insertsort(_ptr(_check_u64(keys)), _ptr(_check_u64(rids)), keys.size)

    def combsort(self, keys, rids):
        self.lib.orc_combsort(_ptr(_check_u64(keys)), _ptr(_check_u64(rids)), keys.size)

    def histogram(self, keys, shift_bits: int, radix_bits: int) -> np.ndarray:
        count = np.zeros(1 << radix_bits, dtype=np.uint64)
        self.lib.orc_histogram(_ptr(_check_u64(keys)), keys.size, _ptr(count),
                               shift_bits, radix_bits)
        return count

    def partition_ip(self, keys, rids, sizes, shift_bits: int, radix_bits: int):
        offsets = np.zeros(1 << radix_bits, dtype=np.uint64)
        self.lib.orc_partition_ip(_ptr(_check_u64(keys)), _ptr(_check_u64(rids)), keys.size,
                                  _ptr(_check_u64(sizes)), _ptr(offsets),
                                  shift_bits, radix_bits)

    def schedule_passes(self, size: int, bits: int):
        rb = np.zeros(8, dtype=np.int8)
        bf = np.zeros(8, dtype=np.int8)
        p = self.lib.orc_schedule_passes(size, bits, _ptr(rb, _i8p), _ptr(bf, _i8p))
        if p < 0:
            return p, [], []
        return p, rb[:p + 1].tolist(), bf[:p + 1].tolist()

    def sort_range(self, keys, rids, bits: int = 58) -> int:
        return self.lib.orc_sort_range(_ptr(_check_u64(keys)), _ptr(_check_u64(rids)),
                                       keys.size, bits)

    def extract_delimiters(self, sorted_sample, parts: int) -> np.ndarray:
        d = np.zeros(parts + 1, dtype=np.uint64)
        d[parts] = np.uint64(0xFFFFFFFFFFFFFFFF)
        self.lib.orc_extract_delimiters(_ptr(_check_u64(sorted_sample)),
                                        sorted_sample.size, _ptr(d))
        return d

    def range_delimiters(self, sorted_sample):
        td = np.zeros(64, dtype=np.uint64)
        rd = np.zeros(128, dtype=np.uint64)
        self.lib.orc_range_delimiters(_ptr(_check_u64(sorted_sample)), sorted_sample.size,
                                      _ptr(td), _ptr(rd))
        return td, rd

    def range_histogram(self, keys, delim):
        ranges = np.zeros(keys.size, dtype=np.uint8)
        count = np.zeros(128, dtype=np.uint64)
        self.lib.orc_range_histogram(_ptr(_check_u64(keys)),
                                     ranges.ctypes.data_as(C.POINTER(C.c_uint8)),
                                     keys.size, _ptr(count), _ptr(_check_u64(delim)))
        return ranges, count

    # -- whole sort --------------------------------------------------------
    def sort(self, keys_list, rids_list, sizes, seed: int = 1) -> list[int]:
        """In place over per-node arrays (arrays longer than sizes[n] give the slack).
        Returns the new per-node sizes."""
        numa = len(keys_list)
        KA = (_u64p * numa)(*[_ptr(_check_u64(k)) for k in keys_list])
        RA = (_u64p * numa)(*[_ptr(_check_u64(r)) for r in rids_list])
        sz = np.array(sizes, dtype=np.uint64)
        cap = np.array([k.size for k in keys_list], dtype=np.uint64)
        rc = self.lib.orc_sort(KA, RA, _ptr(sz), _ptr(cap), numa, seed)
        if rc != 0:
            raise RuntimeError("oracle sort: capacity exceeded or size out of schedule range")
        return [int(x) for x in sz]

    def sort1(self, keys, rids, seed: int = 1):
        self.sort([keys], [rids], [keys.size], seed)

    # -- checks ------------------------------------------------------------
    def check_sorted(self, keys):
        cs = C.c_uint64(0)
        bad = self.lib.orc_check_sorted(_ptr(_check_u64(keys)), keys.size, C.byref(cs))
        return int(bad), int(cs.value)

    def pair_digest(self, keys, rids) -> int:
        return int(self.lib.orc_pair_digest(_ptr(_check_u64(keys)), _ptr(_check_u64(rids)),
                                            keys.size))

    def key_sequence_digest(self, keys) -> int:
        return int(self.lib.orc_key_sequence_digest(_ptr(_check_u64(keys)), keys.size))


class RefLib:
    """The unmodified reference, compiled from /root/reference into oracle/_ref/."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.mamalloc.restype = C.c_void_p
        L.mamalloc.argtypes = [C.c_size_t]
        L.rand64_init.restype = C.c_void_p
        L.rand64_init.argtypes = [C.c_uint64]
        L.rand64_next.restype = C.c_uint64
        L.rand64_next.argtypes = [C.c_void_p]
        L.mulhi.restype = C.c_uint64
        L.mulhi.argtypes = [C.c_uint64, C.c_uint64]
        L.binary_search_64.restype = C.c_uint64
        L.binary_search_64.argtypes = [_u64p, C.c_uint64, C.c_uint64]
        L.insertsort.argtypes = [_u64p, _u64p, C.c_uint64]
        L.combsort.argtypes = [_u64p, _u64p, C.c_uint64]
        L.histogram.argtypes = [_u64p, C.c_uint64, _u64p, C.c_uint8, C.c_uint8]
        L.partition_ip.argtypes = [_u64p, _u64p, C.c_uint64, _u64p, _u64p, C.c_uint8, C.c_uint8]
        L.partition_ip_buf.argtypes = [_u64p, _u64p, C.c_uint64, _u64p, C.c_uint8, C.c_uint8]
        L.schedule_passes.restype = C.c_int
        L.schedule_passes.argtypes = [C.c_uint64, C.c_int8, _i8p, _i8p]
        L.local_radixsort.argtypes = [_u64p, _u64p, C.c_uint64, _i8p, _i8p, C.c_int,
                                      C.POINTER(_u64p), C.POINTER(_u64p)]
        L.extract_delimiters.argtypes = [_u64p, C.c_uint64, _u64p]
        L.range_histogram.argtypes = [_u64p, C.POINTER(C.c_uint8), C.c_uint64, _u64p, _u64p]
        L.sort.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _u64p,
                           C.c_int, C.c_int, C.c_double, C.POINTER(C.c_char_p), _u64p]
        L.check.restype = C.c_uint64

    def aligned(self, count: int) -> np.ndarray:
        """uint64 array from the reference's own 64-byte-aligned allocator (msb_64.c:111)."""
        p = self.lib.mamalloc(max(count, 8) * 8)
        a = np.ctypeslib.as_array(C.cast(p, _u64p), shape=(max(count, 8),))
        return a[:count]

    def rand64(self, seed: int, count: int) -> np.ndarray:
        st = self.lib.rand64_init(seed)
        return np.array([self.lib.rand64_next(st) for _ in range(count)], dtype=np.uint64)

    def schedule_passes(self, size: int, bits: int):
        rb = np.zeros(8, dtype=np.int8)
        bf = np.zeros(8, dtype=np.int8)
        p = self.lib.schedule_passes(size, bits, _ptr(rb, _i8p), _ptr(bf, _i8p))
        return p, rb[:p + 1].tolist(), bf[:p + 1].tolist()

    def local_sort_range(self, keys, rids, bits: int = 58):
        """schedule_passes + suffix sums + local_radixsort, as msb_64.c:2232-2252."""
        rb = np.zeros(8, dtype=np.int8)
        bf = np.zeros(8, dtype=np.int8)
        p = self.lib.schedule_passes(keys.size, bits, _ptr(rb, _i8p), _ptr(bf, _i8p))
        for i in range(p - 1, -1, -1):
            rb[i] += rb[i + 1]
        hist = [self.aligned(4096) for _ in range(5)]
        offs = [self.aligned(4096) for _ in range(5)]
        H = (_u64p * 5)(*[_ptr(h) for h in hist])
        O = (_u64p * 5)(*[_ptr(o) for o in offs])
        self.lib.local_radixsort(_ptr(keys), _ptr(rids), keys.size, _ptr(rb, _i8p),
                                 _ptr(bf, _i8p), 0, H, O)

    def sort(self, keys, rids, n: int, fudge: float, threads: int = 64):
        """Single-node reference sort; keys/rids from aligned(n*fudge+...)."""
        KA = (C.c_void_p * 1)(keys.ctypes.data)
        RA = (C.c_void_p * 1)(rids.ctypes.data)
        SZ = (C.c_uint64 * 1)(n)
        desc = (C.c_char_p * 16)()
        times = (C.c_uint64 * 16)()
        self.lib.sort(KA, RA, SZ, threads, 1, C.c_double(fudge), desc, times)
        phases = {}
        for i in range(16):
            if not desc[i]:
                break
            phases[desc[i].decode().strip().rstrip(":")] = int(times[i])
        return int(SZ[0]), phases


def min_fudge(n: int, threads: int = 64, block_cap: int = 4096, ranges: int = 128) -> float:
    """Smallest fudge the reference accepts for one node (msb_64.c:1574-1578)."""
    blocks = -(-n // block_cap)
    need = 1 + blocks + threads * ranges
    return (need + 2) * block_cap / n
