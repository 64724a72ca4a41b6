"""The N>1 host logic of ShardedSorter on CPU: world_size 2 and 3, gloo backend.

The collectives, the range cut, the count bookkeeping and the capacity error are the
product's own code (inplacemsdradixsort_b200/distributed.py).  The three device steps
(digit histogram, route, local sort) are CUDA kernels in the product and cannot run
here, so this file supplies numpy / oracle stand-ins for them through the `ops` hook
-- test infrastructure only; the GPU versions of the same steps are covered by
tests/test_gpu_distributed.py.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from inputs import make  # noqa: E402


class NumpyOps:
    """Stand-ins for the device steps (same argument meaning as CudaOps)."""

    def __init__(self):
        from oracle.oracle import Oracle
        self.oracle = Oracle()

    def empty(self, count, dtype=None):
        return torch.zeros(int(count), dtype=dtype or torch.int64)

    def from_numpy(self, a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def workspace(self, cap):
        nbytes = 16 * cap + 4096
        return torch.zeros(nbytes, dtype=torch.uint8), nbytes

    @staticmethod
    def _digit(k, shift, bits, origin):
        return (((k >> np.uint64(shift)) - np.uint64(origin)) & np.uint64((1 << bits) - 1)).astype(np.int64)

    def digit_histogram(self, keys, n, shift, bits, origin, out, minmax):
        k = keys[:n].numpy().view(np.uint64)
        d = self._digit(k, shift, bits, origin)
        out[: 1 << bits] = torch.from_numpy(np.bincount(d, minlength=1 << bits).astype(np.int64))
        mm = np.array([k.min() if n else np.uint64((1 << 64) - 1), k.max() if n else 0], dtype=np.uint64)
        minmax.copy_(torch.from_numpy(mm.view(np.int64)))

    def route(self, keys, rids, n, shift, bits, origin, table, world, cursors, out_keys, out_rids):
        k = keys[:n].numpy().view(np.uint64)
        d = self._digit(k, shift, bits, origin)
        dest = table.numpy()[d]
        order = np.argsort(dest, kind="stable")
        starts = cursors.numpy().view(np.uint32)
        counts = np.bincount(dest, minlength=world)
        assert all(int(starts[r]) == int(counts[:r].sum()) for r in range(world))
        out_keys[:n] = keys[:n][torch.from_numpy(order)]
        out_rids[:n] = rids[:n][torch.from_numpy(order)]

    def sort(self, keys, rids, n, ws, ws_bytes, key_lo=0, key_hi=(1 << 64) - 1):
        if n == 0:
            return
        kk = keys[:n].numpy().view(np.uint64)
        assert int(kk.min()) >= key_lo and int(kk.max()) <= key_hi, "received keys outside the rank's range"
        k = np.concatenate([keys[:n].numpy().view(np.uint64), np.zeros(n // 2 + 64, np.uint64)])
        r = np.concatenate([rids[:n].numpy().view(np.uint64), np.zeros(n // 2 + 64, np.uint64)])
        self.oracle.sort([k], [r], [n])
        keys[:n] = torch.from_numpy(k[:n].view(np.int64))
        rids[:n] = torch.from_numpy(r[:n].view(np.int64))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, n, fudge, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from inplacemsdradixsort_b200 import Msb64Error
        from inplacemsdradixsort_b200.distributed import ShardedSorter
        n_local = n + 17 * rank                                   # ragged shards
        keys = make(kind, n_local, seed=100 + rank)
        rids = (np.arange(n_local, dtype=np.uint64) + np.uint64(rank << 40))
        sorter = ShardedSorter(n_local + 64, fudge=fudge, ops=NumpyOps())
        kt = torch.from_numpy(keys.view(np.int64).copy())
        rt = torch.from_numpy(rids.view(np.int64).copy())
        try:
            ok_, or_, cnt = sorter.sort(kt, rt)
        except Msb64Error as e:
            result[rank] = ("error", e.code)
            return
        ordered = sorter.boundaries_ordered(ok_, cnt)
        result[rank] = ("ok", ok_.numpy().view(np.uint64).copy(), or_.numpy().view(np.uint64).copy(),
                        keys, rids, ordered, sorter.last_counts.copy())
    finally:
        dist.destroy_process_group()


def _run(world, kind, n, fudge=1.5):
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kind, n, fudge, result), nprocs=world, join=True)
    return [result[r] for r in range(world)]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["uniform", "low24", "sorted", "midbits", "dup1000"])
def test_sharded_sort_matches_global_sort(world, kind):
    # low24 / midbits: the top 12 bits are equal everywhere, the window moves to the span
    res = _run(world, kind, 20_000, fudge=1.5)
    assert all(r[0] == "ok" for r in res)
    all_k = np.concatenate([r[3] for r in res])
    all_r = np.concatenate([r[4] for r in res])
    out_k = np.concatenate([r[1] for r in res])
    out_r = np.concatenate([r[2] for r in res])
    order = np.lexsort((all_r, all_k))
    assert np.array_equal(out_k, all_k[order]), "concatenated rank outputs != globally sorted keys"
    got = np.lexsort((out_r, out_k))
    assert np.array_equal(out_r[got], all_r[order]), "(key, rid) multiset changed"
    assert all(r[5] for r in res), "rank boundaries out of order"
    counts = res[0][6]
    assert all(np.array_equal(r[6], counts) for r in res), "ranks disagree on the exchange plan"
    assert counts.sum() == all_k.size
    assert [int(c) for c in counts.sum(axis=0)] == [r[1].size for r in res]


def test_uniform_is_balanced():
    res = _run(2, "uniform", 50_000, fudge=1.05)
    sizes = [r[1].size for r in res]
    assert abs(sizes[0] - sizes[1]) < 0.02 * sum(sizes)


def test_capacity_error_like_the_reference_assert():
    # all keys equal: they may not be split, one rank would receive everything
    res = _run(2, "equal", 20_000, fudge=1.25)
    assert all(r[0] == "error" and r[1] == -4 for r in res), "every rank must raise together"


def test_choose_ranges_properties():
    from inplacemsdradixsort_b200.distributed import choose_ranges, exchange_counts
    rng = np.random.default_rng(3)
    for world in (1, 2, 5, 8):
        h = rng.integers(0, 1000, size=4096).astype(np.uint64)
        t = choose_ranges(h, world)
        assert t.dtype == np.uint8 and t.size == 4096
        assert np.all(np.diff(t.astype(np.int64)) >= 0) and t.max() <= world - 1
        per = np.bincount(t, weights=h.astype(np.float64), minlength=world)
        assert per.max() - per.min() <= 2 * 1000 + 1
        hs = rng.integers(0, 50, size=(world, 4096)).astype(np.int64)
        c = exchange_counts(hs, t, world)
        assert np.array_equal(c.sum(axis=1), hs.sum(axis=1))
    assert np.all(choose_ranges(np.zeros(16, np.uint64), 4) == 0)


def test_choose_ranges_hypothesis():
    """Property test of the range cut (host logic, no process group): monotone table, every
    destination index valid, conservation of counts, and balance within the weight of the
    two heaviest bins (bins are never split, so that is the best any cut can promise)."""
    from hypothesis import given, settings, strategies as st
    from inplacemsdradixsort_b200.distributed import choose_ranges, exchange_counts

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 64), st.lists(st.integers(0, 1 << 40), min_size=1, max_size=300),
           st.integers(0, 2 ** 32 - 1))
    def check(world, weights, seed):
        h = np.array(weights, dtype=np.uint64)
        t = choose_ranges(h, world)
        assert t.shape == h.shape and t.dtype == np.uint8
        assert np.all(np.diff(t.astype(np.int64)) >= 0)
        assert int(t.max()) <= world - 1
        per = np.bincount(t, weights=h.astype(np.float64), minlength=world)
        assert per.sum() == float(h.astype(np.float64).sum())
        total = float(h.astype(np.float64).sum())
        if total > 0 and world > 1:
            top2 = float(np.sort(h.astype(np.float64))[-2:].sum())
            assert per.max() <= total / world + top2 + 1e-6 * total
        rng = np.random.default_rng(seed)
        hs = rng.integers(0, 1000, size=(world, h.size)).astype(np.int64)
        c = exchange_counts(hs, t, world)
        assert c.shape == (world, world)
        assert np.array_equal(c.sum(axis=1), hs.sum(axis=1))
        for d in range(world):
            assert int(c[:, d].sum()) == int(hs[:, t == d].sum())

    check()
