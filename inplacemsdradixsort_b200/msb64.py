"""Host-side mirror of the reference's msb_64 interface over the C ABI of libmsb64_b200.so.

Names, argument meaning and failure behaviour follow the reference
(include/msb_64.h:36-40, src/msb_64.c):

    sort(keys, rids, size, threads=64, numa=None, fudge=1.0)   msb_64.c:2261
    mamalloc(count)                                            msb_64.c:111
    check(keys, rids, size, numa, same)                        msb_64.c:2470

plus the device-resident form a GPU pipeline uses (sort_device, DeviceArray).
Everything here only marshals pointers; all work happens in the CUDA library, and
loading fails loudly when that library is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_u64p = C.POINTER(C.c_uint64)
_u64pp = C.POINTER(_u64p)

MSB64_OK = 0
MSB64_MAX_PAIRS = 0xFFFF0000
MSB64_SHARD_HANDLE_BYTES = 192
PHASES = ("histogram", "plan", "scatter", "local_sort", "copy_home", "tail")
ERRORS = {-1: "CUDA", -2: "ARG", -3: "TOO_BIG", -4: "CAPACITY", -5: "NOMEM", -6: "INTERNAL"}

# every symbol include/msb64_b200.h declares (checked by tests without a GPU)
EXPORTS = (
    "sort", "mamalloc", "msb64_b200_sort", "msb64_b200_sort_host",
    "msb64_b200_workspace_bytes", "msb64_b200_sort_device", "msb64_b200_sort_device_range",
    "msb64_b200_get_schedule", "msb64_b200_get_range_schedule",
    "msb64_b200_set_schedule", "msb64_b200_device_count", "msb64_b200_last_error",
    "msb64_b200_launch_count", "msb64_b200_last_stats", "msb64_b200_last_level_times",
    "msb64_b200_host_alloc",
    "msb64_b200_host_free", "msb64_b200_device_alloc", "msb64_b200_device_free",
    "msb64_b200_memcpy_h2d", "msb64_b200_memcpy_d2h", "msb64_b200_memcpy_d2d",
    "msb64_b200_stream_sync", "msb64_b200_fill", "msb64_b200_check",
    "msb64_b200_digit_histogram", "msb64_b200_route", "msb64_b200_route_peer",
    "msb64_b200_ipc_export", "msb64_b200_ipc_open", "msb64_b200_ipc_close",
    "msb64_b200_last_status",
    "msb64_b200_shard_create", "msb64_b200_shard_destroy", "msb64_b200_shard_export",
    "msb64_b200_shard_connect_ipc", "msb64_b200_shard_connect_local", "msb64_b200_shard_slots",
    "msb64_b200_shard_subs", "msb64_b200_shard_histogram", "msb64_b200_shard_hist",
    "msb64_b200_shard_plan", "msb64_b200_shard_plan_host", "msb64_b200_shard_exchange_sort",
    "msb64_b200_shard_count", "msb64_b200_shard_sent", "msb64_b200_shard_sent_direct", "msb64_b200_shard_recv_capacity", "msb64_b200_shard_keys",
    "msb64_b200_shard_rids", "msb64_b200_shard_key_range", "msb64_b200_shard_times",
)


class Msb64Error(RuntimeError):
    def __init__(self, code: int, detail: str):
        super().__init__(f"msb64_b200 error {code} ({ERRORS.get(code, '?')}): {detail}")
        self.code = code


_lib = None


def library_path() -> str:
    # MSB64_B200_LIB: developer override used to compare tuning variants of the library
    return os.environ.get("MSB64_B200_LIB") or _build.LIB_PATH


def load_library() -> C.CDLL:
    """dlopen the CUDA library (never builds, never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: run `python -m inplacemsdradixsort_b200.build` "
            "(needs nvcc); this package has no CPU implementation")
    L = C.CDLL(path)
    L.sort.restype = None
    L.sort.argtypes = [_u64pp, _u64pp, _u64p, C.c_int, C.c_int, C.c_double,
                       C.POINTER(C.c_char_p), _u64p]
    L.mamalloc.restype = C.c_void_p
    L.mamalloc.argtypes = [C.c_size_t]
    L.msb64_b200_sort.restype = C.c_int
    L.msb64_b200_sort.argtypes = L.sort.argtypes
    L.msb64_b200_sort_host.restype = C.c_int
    L.msb64_b200_sort_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.msb64_b200_workspace_bytes.restype = C.c_size_t
    L.msb64_b200_workspace_bytes.argtypes = [C.c_uint64]
    L.msb64_b200_sort_device.restype = C.c_int
    L.msb64_b200_sort_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                         C.c_size_t, C.c_void_p, _u64p]
    L.msb64_b200_sort_device_range.restype = C.c_int
    L.msb64_b200_sort_device_range.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                               C.c_size_t, C.c_void_p, _u64p, C.c_uint64, C.c_uint64]
    L.msb64_b200_get_schedule.restype = C.c_int
    L.msb64_b200_get_schedule.argtypes = [C.c_uint64, C.POINTER(C.c_int)]
    L.msb64_b200_get_range_schedule.restype = C.c_int
    L.msb64_b200_get_range_schedule.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int),
                                                C.POINTER(C.c_int), _u64p]
    L.msb64_b200_set_schedule.restype = C.c_int
    L.msb64_b200_set_schedule.argtypes = [C.POINTER(C.c_int), C.c_int]
    L.msb64_b200_device_count.restype = C.c_int
    L.msb64_b200_last_error.restype = C.c_char_p
    L.msb64_b200_launch_count.restype = C.c_uint64
    L.msb64_b200_last_stats.restype = C.c_int
    L.msb64_b200_last_stats.argtypes = [_u64p, C.c_int]
    L.msb64_b200_last_level_times.restype = C.c_int
    L.msb64_b200_last_level_times.argtypes = [_u64p, C.c_int]
    L.msb64_b200_host_alloc.restype = C.c_void_p
    L.msb64_b200_host_alloc.argtypes = [C.c_size_t]
    L.msb64_b200_host_free.argtypes = [C.c_void_p]
    L.msb64_b200_device_alloc.restype = C.c_void_p
    L.msb64_b200_device_alloc.argtypes = [C.c_size_t]
    L.msb64_b200_device_free.argtypes = [C.c_void_p]
    for name in ("h2d", "d2h", "d2d"):
        f = getattr(L, f"msb64_b200_memcpy_{name}")
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.msb64_b200_stream_sync.restype = C.c_int
    L.msb64_b200_stream_sync.argtypes = [C.c_void_p]
    L.msb64_b200_fill.restype = C.c_int
    L.msb64_b200_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64,
                                  C.c_uint64, C.c_void_p]
    L.msb64_b200_check.restype = C.c_int
    L.msb64_b200_check.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, _u64p, C.c_void_p]
    L.msb64_b200_digit_histogram.restype = C.c_int
    L.msb64_b200_digit_histogram.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
    L.msb64_b200_route.restype = C.c_int
    L.msb64_b200_route.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.msb64_b200_route_peer.restype = C.c_int
    L.msb64_b200_route_peer.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64,
                                        C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.c_void_p]
    L.msb64_b200_ipc_export.restype = C.c_int
    L.msb64_b200_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
    L.msb64_b200_ipc_open.restype = C.c_void_p
    L.msb64_b200_ipc_open.argtypes = [C.c_void_p]
    L.msb64_b200_ipc_close.restype = C.c_int
    L.msb64_b200_ipc_close.argtypes = [C.c_void_p]
    L.msb64_b200_last_status.restype = C.c_int
    L.msb64_b200_last_status.argtypes = [C.c_void_p]
    # section 5: the sort sharded over several GPUs
    L.msb64_b200_shard_create.restype = C.c_void_p
    L.msb64_b200_shard_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_double]
    L.msb64_b200_shard_destroy.restype = None
    L.msb64_b200_shard_destroy.argtypes = [C.c_void_p]
    L.msb64_b200_shard_export.restype = C.c_int
    L.msb64_b200_shard_export.argtypes = [C.c_void_p, C.c_void_p]
    L.msb64_b200_shard_connect_ipc.restype = C.c_int
    L.msb64_b200_shard_connect_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.msb64_b200_shard_connect_local.restype = C.c_int
    L.msb64_b200_shard_connect_local.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.msb64_b200_shard_slots.restype = C.c_int
    L.msb64_b200_shard_subs.restype = C.c_int
    L.msb64_b200_shard_subs.argtypes = [C.c_int]
    L.msb64_b200_shard_histogram.restype = C.c_int
    L.msb64_b200_shard_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    L.msb64_b200_shard_hist.restype = C.c_void_p
    L.msb64_b200_shard_hist.argtypes = [C.c_void_p]
    L.msb64_b200_shard_plan.restype = C.c_int
    L.msb64_b200_shard_plan.argtypes = [C.c_void_p, _u64p, _u64p, C.c_int]
    L.msb64_b200_shard_plan_host.restype = C.c_int
    L.msb64_b200_shard_plan_host.argtypes = [_u64p, C.c_int, _u64p, C.c_int, C.POINTER(C.c_int),
                                             C.POINTER(C.c_int), _u64p, C.c_void_p, _u64p]
    L.msb64_b200_shard_exchange_sort.restype = C.c_int
    L.msb64_b200_shard_exchange_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                                 C.c_void_p, C.c_int]
    for name in ("count", "sent", "sent_direct", "recv_capacity"):
        f = getattr(L, f"msb64_b200_shard_{name}")
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p]
    for name in ("keys", "rids"):
        f = getattr(L, f"msb64_b200_shard_{name}")
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p]
    L.msb64_b200_shard_key_range.restype = C.c_int
    L.msb64_b200_shard_key_range.argtypes = [C.c_void_p, _u64p, _u64p]
    L.msb64_b200_shard_times.restype = C.c_int
    L.msb64_b200_shard_times.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    _lib = L
    return L


def shard_plan_host(hists: np.ndarray, recv_caps, may_retry: bool = True, shift: int = 52, bits: int = 12,
                    origin: int = 0):
    """The plan of the sharded sort on the host (msb64_b200_shard_plan_host, no device needed).

    hists: [world, msb64_b200_shard_slots()] uint64 rows (counts of the digit (key >> shift) -
    origin, then min and max key at [2^bits], [2^bits + 1]).  Returns (rc, shift, bits, origin,
    table, counts): rc 0 = accepted (table[bin] = bucket = destination * subs + sub-range,
    counts[source][bucket]), 1 = histogram again with the returned digit, -4 = CAPACITY."""
    L = load_library()
    hists = np.ascontiguousarray(hists, dtype=np.uint64)
    world = hists.shape[0]
    assert hists.shape[1] == L.msb64_b200_shard_slots()
    subs = L.msb64_b200_shard_subs(world)
    caps = (C.c_uint64 * world)(*[int(x) for x in recv_caps])
    sh, bi, org = C.c_int(shift), C.c_int(bits), C.c_uint64(origin)
    table = np.zeros(2 << 12, dtype=np.uint8)
    counts = np.zeros((world, world * subs), dtype=np.uint64)
    rc = L.msb64_b200_shard_plan_host(hists.ctypes.data_as(_u64p), world, caps, int(may_retry), C.byref(sh),
                                      C.byref(bi), C.byref(org), table.ctypes.data, counts.ctypes.data_as(_u64p))
    return rc, sh.value, bi.value, org.value, table[: 1 << bi.value], counts


def _raise(rc: int) -> None:
    if rc != MSB64_OK:
        raise Msb64Error(rc, load_library().msb64_b200_last_error().decode())


def device_count() -> int:
    return int(load_library().msb64_b200_device_count())


def launch_count() -> int:
    return int(load_library().msb64_b200_launch_count())


# ----------------------------------------------------------------- host memory
def _view(p: int, count: int) -> np.ndarray:
    buf = (C.c_uint64 * max(int(count), 1)).from_address(p)
    return np.frombuffer(buf, dtype=np.uint64)[:count]


def mamalloc(count: int) -> np.ndarray:
    """`count` uint64 slots of 64-byte aligned host memory (msb_64.c:111-115); release
    with mafree() (the C caller uses free(), as with the reference)."""
    p = load_library().mamalloc(max(int(count), 1) * 8)
    if not p:
        raise MemoryError("mamalloc")
    return _view(p, count)


def mafree(arr: np.ndarray) -> None:
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(arr.__array_interface__["data"][0])


def pinned(count: int) -> np.ndarray:
    """`count` uint64 slots of page-locked host memory (fast host<->device copies);
    release with free_pinned()."""
    L = load_library()
    p = L.msb64_b200_host_alloc(max(int(count), 1) * 8)
    if not p:
        raise MemoryError("msb64_b200_host_alloc: " + L.msb64_b200_last_error().decode())
    return _view(p, count)


def free_pinned(arr: np.ndarray) -> None:
    load_library().msb64_b200_host_free(arr.__array_interface__["data"][0])


def _host_ptr(a: np.ndarray):
    if a.dtype != np.uint64 or not a.flags.c_contiguous:
        raise TypeError("keys and rids must be C-contiguous uint64 arrays")
    return a.ctypes.data_as(_u64p)


# ----------------------------------------------------------------- the reference's sort()
def sort(keys, rids, size, threads: int = 64, numa: int | None = None, fudge: float = 1.0):
    """In-place global sort of `numa` host array pairs; see include/msb64_b200.h.

    keys, rids: lists of uint64 arrays (capacity >= size[n] * fudge each); size: list
    of ints, updated in place with the pairs each node holds afterwards.  Returns the
    phase times as {description: microseconds}.  Raises Msb64Error where the
    reference would assert.
    """
    L = load_library()
    if isinstance(keys, np.ndarray):
        keys, rids = [keys], [rids]
    numa = len(keys) if numa is None else numa
    if numa < 1 or len(keys) < numa or len(rids) < numa or len(size) < numa:
        raise Msb64Error(-2, "numa does not match the arrays given")
    for n in range(numa):
        cap = int(size[n] * fudge)
        if keys[n].size < min(cap, size[n]) or rids[n].size < min(cap, size[n]):
            raise Msb64Error(-2, f"array {n} is shorter than size[{n}]")
    KA = (_u64p * numa)(*[_host_ptr(k) for k in keys[:numa]])
    RA = (_u64p * numa)(*[_host_ptr(r) for r in rids[:numa]])
    SZ = (C.c_uint64 * numa)(*[int(s) for s in size[:numa]])
    # capacity the caller really has may be less than size*fudge: keep the C side honest
    eff = fudge
    for n in range(numa):
        if size[n]:
            eff = min(eff, min(keys[n].size, rids[n].size) / size[n])
    desc = (C.c_char_p * 16)()
    times = (C.c_uint64 * 16)()
    _raise(L.msb64_b200_sort(KA, RA, SZ, threads, numa, C.c_double(max(eff, 1.0)), desc, times))
    for n in range(numa):
        size[n] = int(SZ[n])
    phases = {}
    for i in range(16):
        if not desc[i]:
            break
        phases[desc[i].decode().strip().rstrip(":")] = int(times[i])
    return phases


def sort_pairs(keys: np.ndarray, rids: np.ndarray) -> None:
    """One host array pair, in place (numa == 1)."""
    if keys.size != rids.size:
        raise Msb64Error(-2, "keys and rids differ in length")
    _raise(load_library().msb64_b200_sort_host(_host_ptr(keys), _host_ptr(rids), keys.size))


def check(keys, rids, size, numa: int | None = None, same: bool = False) -> int:
    """The reference's check() (msb_64.c:2470-2505): asserts ascending keys inside and
    across nodes (and keys == rids when `same`), returns the wrapping key sum.  Runs on
    the device through msb64_b200_check."""
    if isinstance(keys, np.ndarray):
        keys, rids = [keys], [rids]
    numa = len(keys) if numa is None else numa
    total = 0
    prev_last = None
    for n in range(numa):
        if not size[n]:
            continue
        k = keys[n][:size[n]]
        with DeviceArray(size[n]) as dk:
            dk.upload(k)
            bad, s, _ = dk.check(None)
        if bad:
            raise AssertionError(f"node {n}: {bad} descents")
        if same and not np.array_equal(k, rids[n][:size[n]]):
            raise AssertionError(f"node {n}: keys != rids")
        if prev_last is not None and k[0] < prev_last:
            raise AssertionError(f"node {n} starts below node {n - 1}'s last key")
        prev_last = k[-1]
        total = (total + s) & 0xFFFFFFFFFFFFFFFF
    return total


# ----------------------------------------------------------------- device-resident form
class DeviceArray:
    """n uint64 slots of device memory owned through the C ABI."""

    def __init__(self, count: int):
        self.count = int(count)
        L = load_library()
        self.ptr = L.msb64_b200_device_alloc(max(self.count, 1) * 8)
        if not self.ptr:
            raise Msb64Error(-5, f"device_alloc({self.count * 8} bytes)")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()

    def free(self):
        if self.ptr:
            load_library().msb64_b200_device_free(self.ptr)
            self.ptr = None

    def upload(self, a: np.ndarray, stream=None):
        L = load_library()
        _raise(L.msb64_b200_memcpy_h2d(self.ptr, a.ctypes.data, a.size * 8, stream))
        _raise(L.msb64_b200_stream_sync(stream))

    def download(self, out: np.ndarray | None = None, stream=None) -> np.ndarray:
        L = load_library()
        if out is None:
            out = np.empty(self.count, dtype=np.uint64)
        _raise(L.msb64_b200_memcpy_d2h(out.ctypes.data, self.ptr, out.size * 8, stream))
        _raise(L.msb64_b200_stream_sync(stream))
        return out

    def copy_from(self, other: "DeviceArray", stream=None):
        _raise(load_library().msb64_b200_memcpy_d2d(self.ptr, other.ptr, self.count * 8, stream))

    def check(self, rids: "DeviceArray | None", stream=None):
        out = (C.c_uint64 * 3)()
        _raise(load_library().msb64_b200_check(self.ptr, rids.ptr if rids else None, self.count,
                                               out, stream))
        return int(out[0]), int(out[1]), int(out[2])


def fill(keys: DeviceArray, rids: DeviceArray | None, kind: int = 0, seed: int = 1,
         param: int = 0, stream=None) -> None:
    _raise(load_library().msb64_b200_fill(keys.ptr, rids.ptr if rids else None, keys.count,
                                          kind, seed, param, stream))


def workspace_bytes(n: int) -> int:
    return int(load_library().msb64_b200_workspace_bytes(n))


def sort_device(keys_ptr: int, rids_ptr: int, n: int, workspace_ptr: int | None = None,
                workspace_bytes_: int = 0, stream: int | None = None, timed: bool = False,
                key_range: tuple[int, int] | None = None):
    """Sort n device-resident pairs in place.  Pointers are plain integers (a torch
    tensor's data_ptr(), a DeviceArray's ptr, ...).  Enqueues on `stream` without
    synchronising unless timed=True, which returns {phase: microseconds}.
    key_range=(lo, hi): every key is known to lie in [lo, hi] (msb64_b200_sort_device_range)."""
    L = load_library()
    phase = (C.c_uint64 * len(PHASES))() if timed else None
    if key_range is None:
        _raise(L.msb64_b200_sort_device(keys_ptr, rids_ptr, n, workspace_ptr, workspace_bytes_,
                                        stream, phase))
    else:   # every key lies in [lo, hi]: fewer passes for a narrow range
        _raise(L.msb64_b200_sort_device_range(keys_ptr, rids_ptr, n, workspace_ptr, workspace_bytes_,
                                              stream, phase, int(key_range[0]), int(key_range[1])))
    return dict(zip(PHASES, (int(x) for x in phase))) if timed else None


def sort_tensors(keys, rids, timed: bool = False):
    """torch front end: keys, rids are contiguous CUDA tensors of 8-byte integers;
    sorted in place on the current torch stream."""
    import torch
    for t in (keys, rids):
        if not t.is_cuda or not t.is_contiguous() or t.element_size() != 8:
            raise TypeError("need contiguous CUDA tensors of 8-byte integers")
    if keys.numel() != rids.numel():
        raise Msb64Error(-2, "keys and rids differ in length")
    stream = torch.cuda.current_stream(keys.device).cuda_stream
    with torch.cuda.device(keys.device):
        return sort_device(keys.data_ptr(), rids.data_ptr(), keys.numel(), stream=stream,
                           timed=timed)


def get_schedule(n: int) -> list[int]:
    bits = (C.c_int * 16)()
    k = load_library().msb64_b200_get_schedule(n, bits)
    return [int(bits[i]) for i in range(k)]


def get_range_schedule(n: int, key_lo: int, key_hi: int):
    """(digit widths, shift of the first digit, origin of the first digit) of a sort whose
    keys are known to lie in [key_lo, key_hi] (host logic only, no device needed)."""
    bits = (C.c_int * 32)()
    shift0 = C.c_int()
    origin0 = C.c_uint64()
    k = load_library().msb64_b200_get_range_schedule(n, key_lo, key_hi, bits, C.byref(shift0),
                                                     C.byref(origin0))
    return [int(bits[i]) for i in range(k)], int(shift0.value), int(origin0.value)


def set_schedule(bits: list[int] | None) -> None:
    L = load_library()
    if not bits:
        _raise(L.msb64_b200_set_schedule(None, 0))
    else:
        arr = (C.c_int * len(bits))(*bits)
        _raise(L.msb64_b200_set_schedule(arr, len(bits)))


def last_level_times() -> list[dict]:
    """Per-level device times (microseconds) of the last timed sort_device call."""
    out = (C.c_uint64 * 48)()
    k = load_library().msb64_b200_last_level_times(out, 48)
    return [{"histogram": int(out[i]), "plan": int(out[i + 1]), "scatter": int(out[i + 2])}
            for i in range(0, k, 3)]


def last_stats() -> dict:
    out = (C.c_uint64 * 64)()
    k = load_library().msb64_b200_last_stats(out, 64)
    v = [int(out[i]) for i in range(k)]
    if k < 53:
        return {}
    return {"segments": v[0:16], "tiles": v[16:32], "units": v[32], "copy_tiles": v[33],
            "error": v[34], "degenerate": v[35], "local_pairs": v[36], "moved": v[37:53],
            "hist_keys": v[53] if k > 53 else None}
