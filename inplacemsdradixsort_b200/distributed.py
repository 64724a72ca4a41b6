"""Sort sharded over several B200s: one process per GPU, range partition by key, local sort.

The role of the reference's cross-NUMA-node phase (sample -> range histogram ->
partition into per-node ranges -> every thread sorts its range; msb_64.c:1546-1606,
1674-2198), re-thought for GPUs connected by NVLink/NVSwitch.  Three forms of the exchange:

"pipelined" (the product path; csrc/msb64_shard.cuh has the details)
    1. every rank histograms the top 12 bits of its keys (+ min / max key)   (device kernel)
    2. the histograms are all-gathered                                       (NCCL, or gloo)
    3. every rank computes the same cut of the bin axis into world x 16 buckets -- ascending
       key ranges of near-equal count, 16-32 sub-ranges per destination -- and from it every
       count and offset on every GPU (msb64_b200_shard_plan; a narrow key span gets a
       second histogram round on a window over [min, max])                   (host, C++)
    4. ONE MSD pass at HBM speed groups the rank's pairs by bucket; the first quarter of every
       peer's sub-ranges it stores straight into that peer's receive buffer over NVLink
       (fused compute + exchange), the rest into a local staging buffer       (device kernel)
    5. sub-range by sub-range the copy engines move the staged buckets into the destination
       GPUs' receive buffers over NVLink (CUDA IPC peer memory) and raise a flag behind each
    6. while sub-ranges s+1.. are still travelling, the destination sorts sub-range s with the
       single-GPU sort told its key range: NVLink and HBM work overlap.
"peer"  one kernel routes a tile by destination and stores the runs straight into the peers'
        HBM (fused route + exchange, msb64_b200_route_peer); the local sort starts when
        everything has arrived.  Round 1's form, kept for comparison.
"nccl"  route into a local send buffer + NCCL all-to-all (grouped ncclSend/ncclRecv) -- the
        unfused baseline, and what the gloo tests drive with stand-in device steps.

Afterwards rank r holds the r-th key range in ascending order and every key of rank r
is <= every key of rank r+1 (the contract of the reference's sort() across NUMA
nodes, msb_64.c:2180, include/msb_64.h:36).  Like the reference, the partition has a
capacity factor (`fudge`): a rank that would receive more than capacity * fudge pairs
raises Msb64Error(CAPACITY) on every rank where the reference's assert fires
(msb_64.c:1574-1578).

The device steps go through the C ABI (include/msb64_b200.h, sections 4 and 5) and there is
no CPU implementation of them in this package.  `ops` exists so that the host logic of the
"nccl" form can be exercised by multi-process tests on the gloo backend with stand-in device
steps supplied by the test suite.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import msb64 as _m

DEFAULT_BITS = 12


# --------------------------------------------------------------------- host logic
def choose_ranges(global_hist: np.ndarray, world: int) -> np.ndarray:
    """bin -> destination rank: `world` contiguous bin ranges of near-equal count.

    A bin goes to the rank into whose share its midpoint falls; every rank computes
    this from the same all-gathered histogram and gets the same table.  Bins are never
    split (equal keys never straddle ranks, msb_64.c:1596-1606)."""
    h = np.asarray(global_hist, dtype=np.uint64).astype(np.float64)
    total = float(h.sum())
    if total == 0 or world == 1:
        return np.zeros(h.size, dtype=np.uint8)
    mid = np.cumsum(h) - h / 2.0
    dest = np.minimum((mid * world / total).astype(np.int64), world - 1)
    return np.maximum.accumulate(dest).astype(np.uint8)       # monotone by construction; keep it exact


def exchange_counts(hists: np.ndarray, table: np.ndarray, world: int) -> np.ndarray:
    """counts[src][dst] = pairs rank `src` sends to rank `dst` under `table`."""
    counts = np.zeros((world, world), dtype=np.int64)
    for dst in range(world):
        sel = table == dst
        if sel.any():
            counts[:, dst] = hists[:, sel].sum(axis=1)
    return counts


# --------------------------------------------------------------------- device steps
class _RawCuda:
    """__cuda_array_interface__ view of `count` int64 slots at a raw device pointer."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<i8",
                                         "data": (int(ptr), False), "version": 3, "strides": None}


class CudaOps:
    """The device steps on torch CUDA tensors through libmsb64_b200.so."""
    supports_peer = True

    def __init__(self, device):
        import torch
        self.torch = torch
        self.device = device
        self.lib = _m.load_library()
        if not torch.cuda.is_available() or _m.device_count() < 1:
            raise _m.Msb64Error(-1, "ShardedSorter needs a CUDA device; there is no CPU path")

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def empty(self, count, dtype=None):
        return self.torch.empty(int(count), dtype=dtype or self.torch.int64, device=self.device)

    def from_numpy(self, a):
        return self.torch.from_numpy(a).to(self.device, non_blocking=False)

    def workspace(self, cap):
        nbytes = _m.workspace_bytes(cap)
        return self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device), nbytes

    def digit_histogram(self, keys, n, shift, bits, origin, out, minmax):
        """out[0 : 2^bits] = counts of ((key >> shift) - origin) & mask; minmax[0:2] = min, max key."""
        _m._raise(self.lib.msb64_b200_digit_histogram(keys.data_ptr(), n, shift, bits, origin,
                                                      out.data_ptr(), minmax.data_ptr(), self._stream()))

    def route(self, keys, rids, n, shift, bits, origin, table, world, cursors, out_keys, out_rids):
        _m._raise(self.lib.msb64_b200_route(keys.data_ptr(), rids.data_ptr(), n, shift, bits, origin,
                                            table.data_ptr(), world, cursors.data_ptr(),
                                            out_keys.data_ptr(), out_rids.data_ptr(), self._stream()))

    # -- peer-memory exchange
    def alloc_exportable(self, count):
        """cudaMalloc'ed (IPC-exportable) int64 array: (raw pointer, torch view)."""
        with self.torch.cuda.device(self.device):
            ptr = self.lib.msb64_b200_device_alloc(max(int(count), 1) * 8)
        if not ptr:
            raise _m.Msb64Error(-5, f"device_alloc({count * 8} bytes)")
        return ptr, self.torch.as_tensor(_RawCuda(ptr, count), device=self.device)

    def free_exportable(self, ptr):
        self.lib.msb64_b200_device_free(ptr)

    def ipc_export(self, ptr) -> bytes:
        buf = C.create_string_buffer(64)
        _m._raise(self.lib.msb64_b200_ipc_export(ptr, buf))
        return buf.raw

    def ipc_open(self, handle: bytes):
        with self.torch.cuda.device(self.device):
            return self.lib.msb64_b200_ipc_open(C.create_string_buffer(handle, 64))

    def ipc_close(self, ptr):
        self.lib.msb64_b200_ipc_close(ptr)

    def route_peer(self, keys, rids, n, shift, bits, origin, table, world, cursors, key_ptrs, rid_ptrs):
        kp = (C.c_void_p * world)(*key_ptrs)
        rp = (C.c_void_p * world)(*rid_ptrs)
        _m._raise(self.lib.msb64_b200_route_peer(keys.data_ptr(), rids.data_ptr(), n, shift, bits, origin,
                                                 table.data_ptr(), world, cursors.data_ptr(), kp, rp,
                                                 self._stream()))

    def sort(self, keys, rids, n, ws, ws_bytes, key_lo=0, key_hi=(1 << 64) - 1):
        _m._raise(self.lib.msb64_b200_sort_device_range(keys.data_ptr(), rids.data_ptr(), n, ws.data_ptr(),
                                                        ws_bytes, self._stream(), None, key_lo, key_hi))

    # -- pipelined form: one msb64_b200_shard per rank (include/msb64_b200.h section 5)
    supports_pipelined = True

    def shard_create(self, rank, world, capacity, fudge):
        with self.torch.cuda.device(self.device):
            h = self.lib.msb64_b200_shard_create(rank, world, capacity, C.c_double(fudge))
        if not h:
            raise _m.Msb64Error(-5, self.lib.msb64_b200_last_error().decode())
        return h

    def view(self, ptr, count):
        """torch view of `count` int64 slots of library-owned device memory."""
        return self.torch.as_tensor(_RawCuda(ptr, count), device=self.device)


# --------------------------------------------------------------------- the sorter
class ShardedSorter:
    """Reusable buffers + the steps above for `capacity` local pairs per rank.

    exchange: "pipelined" | "peer" | "nccl" | "auto" (pipelined where the device steps are the
    CUDA library's, else nccl).  The collectives run on `group`; with a gloo group (two ranks
    sharing one GPU in the tests, or a CPU control plane) they go through host memory."""

    def __init__(self, capacity: int, device=None, fudge: float = 1.125, bits: int = DEFAULT_BITS,
                 group=None, ops=None, exchange: str = "auto"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.host_collectives = dist.is_initialized() and dist.get_backend(group) == "gloo"
        if self.world > 64:
            raise _m.Msb64Error(-2, "at most 64 ranks (msb64_b200_route destinations)")
        self.bits = int(bits)
        self.shift = 64 - self.bits
        self.capacity = int(capacity)
        self.fudge = float(fudge)
        if not (self.fudge >= 1.0):
            raise _m.Msb64Error(-2, "fudge must be >= 1.0")
        self.recv_cap = int(self.capacity * self.fudge) + 2
        if self.recv_cap > _m.MSB64_MAX_PAIRS:
            raise _m.Msb64Error(-3, "more than MSB64_MAX_PAIRS pairs per GPU")
        self.ops = ops if ops is not None else CudaOps(device)
        o = self.ops
        if not 4 <= self.bits <= 12:
            raise _m.Msb64Error(-2, "bits must be 4..12")
        if exchange not in ("auto", "pipelined", "peer", "nccl"):
            raise _m.Msb64Error(-2, "exchange must be 'auto', 'pipelined', 'peer' or 'nccl'")
        self.exchange = "nccl"
        self._own = self._peer_keys = self._peer_rids = None
        self._shard = None
        self.last_counts = None
        self.last_times = None
        pipelined_ok = getattr(o, "supports_pipelined", False)
        if exchange == "pipelined" and not pipelined_ok:
            raise _m.Msb64Error(-1, "the pipelined exchange needs the CUDA library's device steps")
        if exchange in ("auto", "pipelined") and pipelined_ok:
            self._setup_pipelined(exchange == "pipelined")
        if self.exchange != "pipelined":
            self.slots = (2 << self.bits) + 2              # room for the (bits+1)-bit window + min, max
            self.hist = o.empty(self.slots)
            self.all_hist = o.empty(self.world * self.slots)
            if self.world > 1 and exchange == "peer" and getattr(o, "supports_peer", False):
                self._setup_peer(True)
            if self.exchange == "nccl":
                self.recv_keys = o.empty(self.recv_cap)
                self.recv_rids = o.empty(self.recv_cap)
            # the sort's scratch doubles as the send buffer: the exchange is over before the
            # local sort starts, and the sort treats its workspace as uninitialised
            self.ws, self.ws_bytes = o.workspace(self.recv_cap)
            words = self.ws[: (self.ws.numel() // 8) * 8].view(torch.int64)
            assert words.numel() >= 2 * self.capacity, "workspace smaller than the send buffers"
            self.send_keys = words[: self.capacity]
            self.send_rids = words[self.capacity: 2 * self.capacity]
        # every rank knows every rank's receive capacity, so that an overflow is raised on
        # all ranks together (nobody is left waiting in the exchange)
        self.recv_caps = self._all_gather_host(np.array([self.recv_cap], dtype=np.int64)).reshape(-1)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- collectives: rows of 8-byte words from every rank, on the host
    def _all_gather_host(self, row: np.ndarray, device_row=None) -> np.ndarray:
        """[world, len(row)] int64.  device_row: the same row as a device tensor (saves the
        host round trip in front of an NCCL all-gather)."""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return (row if device_row is None else device_row.cpu().numpy()).reshape(1, -1).copy()
        if self.host_collectives:
            mine = torch.from_numpy(np.ascontiguousarray(row if device_row is None else device_row.cpu().numpy()))
            out = torch.empty(self.world * mine.numel(), dtype=mine.dtype)
            dist.all_gather_into_tensor(out, mine, group=self.group)
            return out.numpy().reshape(self.world, -1)
        mine = device_row if device_row is not None else self.ops.from_numpy(np.ascontiguousarray(row))
        out = self.ops.empty(self.world * mine.numel(), dtype=mine.dtype)
        dist.all_gather_into_tensor(out, mine, group=self.group)
        return out.cpu().numpy().reshape(self.world, -1)             # synchronises

    def _all_agree(self, ok: bool) -> bool:
        flags = self._all_gather_host(np.array([int(ok)], dtype=np.int64))
        return bool(flags.min())

    # -- pipelined exchange: one msb64_b200_shard, the peers' buffers mapped through CUDA IPC
    def _setup_pipelined(self, required: bool):
        o, lib = self.ops, self.ops.lib
        ok, shard, detail = True, None, ""
        try:
            shard = o.shard_create(self.rank, self.world, self.capacity, self.fudge)
        except _m.Msb64Error as e:
            ok, detail = False, str(e)
        handles = np.zeros(_m.MSB64_SHARD_HANDLE_BYTES // 8, dtype=np.int64)
        if ok and self.world > 1:
            buf = C.create_string_buffer(_m.MSB64_SHARD_HANDLE_BYTES)
            if lib.msb64_b200_shard_export(shard, buf) == 0:
                handles = np.frombuffer(buf.raw, dtype=np.int64).copy()
            else:
                ok, detail = False, lib.msb64_b200_last_error().decode()
        # every rank takes part in every collective, whatever happened to it so far
        allh = self._all_gather_host(handles)
        if not self._all_agree(ok):
            ok = False
        if ok and self.world > 1:
            raw = np.ascontiguousarray(allh).tobytes()
            if lib.msb64_b200_shard_connect_ipc(shard, C.create_string_buffer(raw, len(raw))) != 0:
                ok, detail = False, lib.msb64_b200_last_error().decode()
        if self.world > 1 and not self._all_agree(ok):
            ok = False
        if not ok:
            if shard:
                lib.msb64_b200_shard_destroy(shard)
            if required:
                raise _m.Msb64Error(-1, "pipelined exchange unavailable on some rank (CUDA IPC / peer access / "
                                        "memory): " + detail)
            return
        self._shard = shard
        self.exchange = "pipelined"
        self.slots = int(lib.msb64_b200_shard_slots())
        self.recv_cap = int(lib.msb64_b200_shard_recv_capacity(shard))
        self.hist = o.view(lib.msb64_b200_shard_hist(shard), self.slots)
        self.recv_keys = o.view(lib.msb64_b200_shard_keys(shard), self.recv_cap)
        self.recv_rids = o.view(lib.msb64_b200_shard_rids(shard), self.recv_cap)

    def _sort_pipelined(self, keys, rids, n, timed):
        torch, lib, shard = self.torch, self.ops.lib, self._shard
        stream = self.ops._stream()
        ev = None
        if timed:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        caps = (C.c_uint64 * self.world)(*[int(x) for x in self.recv_caps])
        rc, reset = 1, 1
        while rc == 1:
            with torch.cuda.device(self.ops.device):
                _m._raise(lib.msb64_b200_shard_histogram(shard, keys.data_ptr(), n, reset, stream))
            # the all-gather also tells every rank that all of them are inside this call: nobody
            # is still reading the receive buffers the peers are about to write into
            allh = np.ascontiguousarray(self._all_gather_host(None, device_row=self.hist)).view(np.uint64)
            rc = lib.msb64_b200_shard_plan(shard, allh.ctypes.data_as(C.POINTER(C.c_uint64)), caps, reset)
            reset = 0
        _m._raise(rc)                                    # CAPACITY: the same verdict on every rank
        if ev:
            ev[1].record()
        with torch.cuda.device(self.ops.device):
            _m._raise(lib.msb64_b200_shard_exchange_sort(shard, keys.data_ptr(), rids.data_ptr(), n, stream,
                                                         1 if timed else 0))
        total = int(lib.msb64_b200_shard_count(shard))
        if ev:
            ev[2].record()
            torch.cuda.synchronize()
            ms = (C.c_double * 5)()
            _m._raise(lib.msb64_b200_shard_times(shard, ms))
            self.last_times = {"plan": ev[0].elapsed_time(ev[1]), "route": ms[0], "first_wait": ms[1],
                               "sort": ms[2], "exchange": ms[3], "step_device": ms[4],
                               "total": ev[0].elapsed_time(ev[2]), "pairs_received": total,
                               "pairs_sent_to_peers": int(lib.msb64_b200_shard_sent(shard)),
                               "pairs_stored_by_route": int(lib.msb64_b200_shard_sent_direct(shard))}
        return self.recv_keys[:total], self.recv_rids[:total], total

    # -- peer-memory exchange: map every rank's receive buffers into this process
    def _setup_peer(self, required: bool):
        o = self.ops
        ok, opened = True, []
        kp = rp = 0
        handles = np.zeros(16, dtype=np.int64)
        try:
            kp, self.recv_keys = o.alloc_exportable(self.recv_cap)
            rp, self.recv_rids = o.alloc_exportable(self.recv_cap)
            self._own = (kp, rp)
            handles = np.frombuffer(o.ipc_export(kp) + o.ipc_export(rp), dtype=np.int64).copy()
        except _m.Msb64Error:
            ok = False
        # every rank takes part in every collective, whatever happened to it so far
        h = np.ascontiguousarray(self._all_gather_host(handles)).tobytes()
        ok = self._all_agree(ok)
        keys_p, rids_p = [0] * self.world, [0] * self.world
        if ok:
            for r in range(self.world):
                if r == self.rank:
                    keys_p[r], rids_p[r] = kp, rp
                    continue
                a = o.ipc_open(h[128 * r: 128 * r + 64])
                b = o.ipc_open(h[128 * r + 64: 128 * r + 128])
                opened += [x for x in (a, b) if x]
                if not a or not b:
                    ok = False
                    break
                keys_p[r], rids_p[r] = a, b
        if self._all_agree(ok):
            self.exchange = "peer"
            self._peer_keys, self._peer_rids, self._opened = keys_p, rids_p, opened
            return
        for x in opened:
            o.ipc_close(x)
        if self._own:
            self.recv_keys = self.recv_rids = None
            for x in self._own:
                o.free_exportable(x)
            self._own = None
        if required:
            raise _m.Msb64Error(-1, "peer-memory exchange unavailable (CUDA IPC / peer access failed)")

    def close(self):
        """Unmap the peers' buffers and release the exported ones (collective: every rank
        must have finished using the sorter)."""
        if self._shard is not None or (self.exchange == "peer" and self._own):
            if self.torch.cuda.is_available():
                self.torch.cuda.synchronize()
            if self.world > 1:
                self._all_agree(True)                         # a barrier on whatever backend the group has
        if self._shard is not None:
            self.hist = self.recv_keys = self.recv_rids = None
            self.ops.lib.msb64_b200_shard_destroy(self._shard)
            self._shard = None
        if self.exchange == "peer" and self._own:
            for x in self._opened:
                self.ops.ipc_close(x)
            self.recv_keys = self.recv_rids = None
            for x in self._own:
                self.ops.free_exportable(x)
            self._own = None

    def __del__(self):
        # single-rank sorters own nothing a peer has mapped: release without a collective
        try:
            if self.world == 1 and self._shard is not None:
                self.close()
        except Exception:
            pass

    # -- steps 1-3 ("peer" and "nccl" forms)
    def _histograms(self, keys, n, shift, bits, origin):
        """Per-rank histograms [world, 2^bits] of the given digit and the global min / max key."""
        nb = 1 << bits
        self.ops.digit_histogram(keys, n, shift, bits, origin, self.hist, self.hist[nb: nb + 2])
        h = self._all_gather_host(None, device_row=self.hist)
        mm = np.ascontiguousarray(h[:, nb: nb + 2]).view(np.uint64)
        have = mm[:, 0] <= mm[:, 1]                                       # ranks that hold any key
        gmin = int(mm[have, 0].min()) if have.any() else 0
        gmax = int(mm[have, 1].max()) if have.any() else 0
        return h[:, :nb], gmin, gmax

    def plan(self, keys, n):
        """Returns (shift, bits, origin, table, counts): digit = ((key >> shift) - origin) &
        (2^bits - 1), table[digit] = destination rank, counts[src][dst] = pairs to move."""
        shift, bits, origin = self.shift, self.bits, 0
        h, gmin, gmax = self._histograms(keys, n, shift, bits, origin)
        table = choose_ranges(h.sum(axis=0), self.world)
        counts = exchange_counts(h, table, self.world)
        if self.world > 1 and np.any(counts.sum(axis=0) > self.recv_caps):
            # the top bits do not separate the keys: put the window on their real span
            width = (gmax - gmin).bit_length()
            shift2 = max(width - self.bits, 0)
            if shift2 < shift:
                shift, bits, origin = shift2, self.bits + 1, gmin >> shift2   # digits 0 .. 2^bits (one past: bits + 1)
                h, _, _ = self._histograms(keys, n, shift, bits, origin)
                table = choose_ranges(h.sum(axis=0), self.world)
                counts = exchange_counts(h, table, self.world)
        return shift, bits, origin, table, counts

    def sort(self, keys, rids, n: int | None = None, timed: bool = False):
        """keys, rids: this rank's pairs (8-byte integer tensors on the sorter's device).
        Returns (keys, rids, count) of this rank's key range, sorted; the tensors are
        views of the sorter's receive buffers, valid until the next call.
        timed=True (CUDA ops only) records device times of the steps in self.last_times
        (milliseconds) and synchronises."""
        torch, dist = self.torch, self.dist
        n = keys.numel() if n is None else int(n)
        if n > self.capacity:
            raise _m.Msb64Error(-2, f"{n} pairs exceed the sorter's capacity {self.capacity}")
        if self.exchange == "pipelined":
            return self._sort_pipelined(keys, rids, n, timed)
        ev = None
        if timed:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
        shift, bits, origin, table, counts = self.plan(keys, n)
        if ev:
            ev[1].record()
        self.last_counts = counts
        send = counts[self.rank]
        recv = counts[:, self.rank]
        total = int(recv.sum())
        over = np.nonzero(counts.sum(axis=0) > self.recv_caps)[0]
        if over.size:                                     # same verdict on every rank
            r = int(over[0])
            raise _m.Msb64Error(-4, f"rank {r} would receive {int(counts[:, r].sum())} pairs, more "
                                    f"than its capacity * fudge = {int(self.recv_caps[r])}")
        if self.world == 1:
            # nothing to exchange: the receive buffer gets the pairs, the sort runs on it
            self.recv_keys[:n].copy_(keys[:n])
            self.recv_rids[:n].copy_(rids[:n])
        elif self.exchange == "peer":
            # steps 4 + 5 fused: this source's slice of destination d's buffer starts after
            # the pairs of the lower-ranked sources.  The histogram all-gather above already
            # told us that every rank has entered this call (nobody still reads its buffer).
            starts = counts[: self.rank].sum(axis=0).astype(np.uint32)
            cursors = self.ops.from_numpy(starts.view(np.int32))
            table_d = self.ops.from_numpy(table)
            self.ops.route_peer(keys, rids, n, shift, bits, origin, table_d, self.world, cursors,
                                self._peer_keys, self._peer_rids)
            if ev:
                ev[2].record()
            torch.cuda.current_stream().synchronize()          # this rank's stores have left
            self._all_agree(True)                              # ... and so have everybody's
        else:
            starts = np.concatenate([[0], np.cumsum(send)[:-1]]).astype(np.uint32)
            cursors = self.ops.from_numpy(starts.view(np.int32))
            table_d = self.ops.from_numpy(table)
            self.ops.route(keys, rids, n, shift, bits, origin, table_d, self.world, cursors,
                           self.send_keys, self.send_rids)
            ins, outs = [int(x) for x in send], [int(x) for x in recv]
            dist.all_to_all_single(self.recv_keys[:total], self.send_keys[:n], outs, ins, group=self.group)
            dist.all_to_all_single(self.recv_rids[:total], self.send_rids[:n], outs, ins, group=self.group)
        # this rank's keys lie in its bin range: the local sort spends no pass on the bits
        # the range partition already fixed
        mine = np.nonzero(table == self.rank)[0]
        if mine.size:
            key_lo = min((origin + int(mine[0])) << shift, (1 << 64) - 1)
            key_hi = min(((origin + int(mine[-1]) + 1) << shift) - 1, (1 << 64) - 1)
        else:
            key_lo, key_hi = 0, (1 << 64) - 1
        if ev:
            if self.exchange != "peer" or self.world == 1:
                ev[2].record()
            ev[3].record()
        self.ops.sort(self.recv_keys, self.recv_rids, total, self.ws, self.ws_bytes, key_lo, key_hi)
        if ev:
            ev[4].record()
            torch.cuda.synchronize()
            self.last_times = {"plan": ev[0].elapsed_time(ev[1]), "exchange": ev[1].elapsed_time(ev[2]),
                               "barrier": ev[2].elapsed_time(ev[3]), "local_sort": ev[3].elapsed_time(ev[4]),
                               "pairs_sent_to_peers": int(send.sum() - send[self.rank]),
                               "pairs_received": total}
        return self.recv_keys[:total], self.recv_rids[:total], total

    # -- acceptance across ranks (the cross-node half of check(), msb_64.c:2485-2495)
    def boundaries_ordered(self, out_keys, out_n: int) -> bool:
        """True on every rank iff every rank's last key <= the next non-empty rank's first."""
        if self.world == 1:
            return True
        edge = np.zeros(3, dtype=np.int64)
        if out_n:
            ends = out_keys[[0, out_n - 1]].cpu().numpy()
            edge[0], edge[1], edge[2] = ends[0], ends[1], 1
        e = self._all_gather_host(edge)
        first, last = np.ascontiguousarray(e[:, 0]).view(np.uint64), np.ascontiguousarray(e[:, 1]).view(np.uint64)
        prev = None
        for r in range(self.world):
            if not e[r, 2]:
                continue
            if prev is not None and first[r] < prev:
                return False
            prev = last[r]
        return True
