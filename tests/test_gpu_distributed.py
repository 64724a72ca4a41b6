"""The device steps of the multi-GPU range partition (include/msb64_b200.h section 4) and
ShardedSorter on real GPUs: one GPU always, two GPUs over NCCL when the box has them."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from inputs import make  # noqa: E402

pytestmark = pytest.mark.gpu


def _dev(gpu, a):
    d = gpu.DeviceArray(a.size)
    if a.size:
        d.upload(np.ascontiguousarray(a))
    return d


@pytest.mark.parametrize("kind", ["uniform", "low24", "skew", "sorted"])
@pytest.mark.parametrize("n,bits,shift", [(0, 8, 56), (1, 12, 52), (100_003, 12, 52), (1 << 20, 10, 20),
                                          (300_001, 1, 63)])
def test_digit_histogram(gpu, kind, n, bits, shift):
    lib = gpu.load_library()
    keys = make(kind, n, seed=5)
    dk = _dev(gpu, keys)
    dh = gpu.DeviceArray((1 << bits) + 2)
    origin = int(keys.min() >> np.uint64(shift)) if n and kind == "skew" else 0
    assert lib.msb64_b200_digit_histogram(dk.ptr, n, shift, bits, origin, dh.ptr,
                                          dh.ptr + 8 * (1 << bits), None) == 0
    got = dh.download()
    digit = ((keys >> np.uint64(shift)) - np.uint64(origin)) & np.uint64((1 << bits) - 1)
    want = np.bincount(digit.astype(np.int64), minlength=1 << bits).astype(np.uint64)
    assert np.array_equal(got[: 1 << bits], want)
    if n:
        assert (int(got[1 << bits]), int(got[(1 << bits) + 1])) == (int(keys.min()), int(keys.max()))
    else:
        assert int(got[1 << bits]) > int(got[(1 << bits) + 1])            # "no keys": min > max


@pytest.mark.parametrize("kind", ["uniform", "skew", "dup16", "sorted"])
@pytest.mark.parametrize("n,ndest", [(1, 2), (4097, 3), (100_003, 8), (1 << 20, 64), (300_001, 1)])
def test_route_groups_by_destination(gpu, kind, n, ndest):
    lib = gpu.load_library()
    bits, shift = 12, 52
    keys = make(kind, n, seed=9)
    rids = np.arange(n, dtype=np.uint64) * np.uint64(7) + np.uint64(1)
    digit = ((keys >> np.uint64(shift)) & np.uint64((1 << bits) - 1)).astype(np.int64)
    table = (np.arange(1 << bits) * ndest // (1 << bits)).astype(np.uint8)      # contiguous bin ranges
    dest = table[digit]
    counts = np.bincount(dest, minlength=ndest)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint32)
    dk, dr = _dev(gpu, keys), _dev(gpu, rids)
    ok_, or_ = gpu.DeviceArray(n), gpu.DeviceArray(n)
    dt = gpu.DeviceArray((table.size + 7) // 8)
    lib.msb64_b200_memcpy_h2d(dt.ptr, table.ctypes.data, table.size, None)
    dc = gpu.DeviceArray((ndest + 1) // 2 + 1)
    lib.msb64_b200_memcpy_h2d(dc.ptr, starts.ctypes.data, starts.size * 4, None)
    lib.msb64_b200_stream_sync(None)
    assert lib.msb64_b200_route(dk.ptr, dr.ptr, n, shift, bits, 0, dt.ptr, ndest, dc.ptr, ok_.ptr, or_.ptr, None) == 0
    gk, gr = ok_.download(), or_.download()
    cur = np.zeros(dc.count, dtype=np.uint64)
    dc.download(cur)
    ends = cur.view(np.uint32)[:ndest]
    assert np.array_equal(ends.astype(np.int64), (starts + counts).astype(np.int64)), "cursors not advanced by the counts"
    for d in range(ndest):
        lo, hi = int(starts[d]), int(starts[d] + counts[d])
        sel = dest == d
        want = np.lexsort((rids[sel], keys[sel]))
        got = np.lexsort((gr[lo:hi], gk[lo:hi]))
        assert np.array_equal(gk[lo:hi][got], keys[sel][want]), f"destination {d}: keys differ"
        assert np.array_equal(gr[lo:hi][got], rids[sel][want]), f"destination {d}: rids differ"


def _oracle_sorted(oracle, keys, rids):
    n = keys.size
    wk = np.concatenate([keys, np.zeros(n // 2 + 64, np.uint64)])
    wr = np.concatenate([rids, np.zeros(n // 2 + 64, np.uint64)])
    oracle.sort([wk], [wr], [n])
    return wk[:n], wr[:n]


@pytest.mark.parametrize("exchange", ["pipelined", "nccl"])
@pytest.mark.parametrize("kind", ["uniform", "low24", "dup16", "sorted", "midbits", "skew", "equal", "outlier"])
def test_sharded_sorter_single_gpu(gpu, oracle, kind, exchange):
    """world = 1: the pipelined form still runs its bucket pass and the sub-range sorts."""
    import torch
    from inplacemsdradixsort_b200.distributed import ShardedSorter
    n = 200_003
    keys = make(kind, n, seed=3)
    rids = np.arange(n, dtype=np.uint64)
    dev = torch.device("cuda", 0)
    with ShardedSorter(n, dev, exchange=exchange) as s:
        assert s.exchange == exchange
        kt = torch.from_numpy(keys.view(np.int64)).to(dev)
        rt = torch.from_numpy(rids.view(np.int64)).to(dev)
        for _ in range(2):                                   # twice through the same buffers
            ok_, or_, cnt = s.sort(kt, rt)
            torch.cuda.synchronize()
            assert cnt == n
            gk = ok_.cpu().numpy().view(np.uint64)
            gr = or_.cpu().numpy().view(np.uint64)
            wk, wr = _oracle_sorted(oracle, keys, rids)
            assert np.array_equal(gk, wk)
            assert np.array_equal(gr[np.lexsort((gr, gk))], wr[np.lexsort((wr, wk))])
            assert s.boundaries_ordered(ok_, cnt)
        assert gpu.load_library().msb64_b200_last_status(None) == 0


@pytest.mark.parametrize("n", [0, 1, 2, 4097, 1 << 20, 3_000_001])
def test_pipelined_sizes_single_gpu(gpu, n):
    import torch
    from inplacemsdradixsort_b200.distributed import ShardedSorter
    dev = torch.device("cuda", 0)
    keys = make("uniform", n, seed=11)
    rids = np.arange(n, dtype=np.uint64)
    with ShardedSorter(max(n, 1), dev, exchange="pipelined") as s:
        ok_, or_, cnt = s.sort(torch.from_numpy(keys.view(np.int64)).to(dev),
                               torch.from_numpy(rids.view(np.int64)).to(dev), n=n)
        torch.cuda.synchronize()
        assert cnt == n
        gk = ok_.cpu().numpy().view(np.uint64)
        gr = or_.cpu().numpy().view(np.uint64)
        order = np.lexsort((rids, keys))
        assert np.array_equal(gk, keys[order])
        assert np.array_equal(gr[np.lexsort((gr, gk))], rids[order])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, kind, n, exchange, result):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from inplacemsdradixsort_b200.distributed import ShardedSorter
        n_local = n + 1000 * rank
        keys = make(kind, n_local, seed=40 + rank)
        rids = np.arange(n_local, dtype=np.uint64) + np.uint64(rank << 40)
        s = ShardedSorter(n_local, dev, fudge=1.3, exchange=exchange)
        assert s.exchange == exchange
        ok_, or_, cnt = s.sort(torch.from_numpy(keys.view(np.int64)).to(dev),
                               torch.from_numpy(rids.view(np.int64)).to(dev))
        ordered = s.boundaries_ordered(ok_, cnt)
        torch.cuda.synchronize()
        result[rank] = (ok_.cpu().numpy().view(np.uint64).copy(), or_.cpu().numpy().view(np.uint64).copy(),
                        keys, rids, ordered)
        # a second sort through the same buffers (peers write into them again)
        keys2 = make(kind, n_local, seed=400 + rank)
        ok2, or2, cnt2 = s.sort(torch.from_numpy(keys2.view(np.int64)).to(dev),
                                torch.from_numpy(rids.view(np.int64)).to(dev))
        assert s.boundaries_ordered(ok2, cnt2)
        k2 = ok2.cpu().numpy().view(np.uint64)
        assert bool(np.all(k2[:-1] <= k2[1:]))
        s.close()
    finally:
        dist.destroy_process_group()


def _shared_gpu_worker(rank, world, port, kind, n, exchange, result):
    """`world` processes on cuda:0: collectives over gloo (host), the exchange over CUDA IPC
    mappings of the other processes' buffers on the same device -- the N > 1 code path of the
    product on a box with a single GPU."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from inplacemsdradixsort_b200.distributed import ShardedSorter
        n_local = n + 1000 * rank
        s = ShardedSorter(n_local, dev, fudge=1.3, exchange=exchange)
        assert s.exchange == exchange
        outs = []
        for it in range(2):                                  # the peers write into the same buffers again
            keys = make(kind, n_local, seed=40 + rank + 100 * it)
            rids = np.arange(n_local, dtype=np.uint64) + np.uint64(rank << 40)
            ok_, or_, cnt = s.sort(torch.from_numpy(keys.view(np.int64)).to(dev),
                                   torch.from_numpy(rids.view(np.int64)).to(dev))
            ordered = s.boundaries_ordered(ok_, cnt)
            torch.cuda.synchronize()
            outs.append((ok_.cpu().numpy().view(np.uint64).copy(), or_.cpu().numpy().view(np.uint64).copy(),
                         keys, rids, ordered))
        result[rank] = outs
        s.close()
    finally:
        dist.destroy_process_group()


def _check_global(res, it, oracle=None):
    all_k = np.concatenate([r[it][2] for r in res])
    all_r = np.concatenate([r[it][3] for r in res])
    out_k = np.concatenate([r[it][0] for r in res])
    out_r = np.concatenate([r[it][1] for r in res])
    order = np.lexsort((all_r, all_k))
    assert np.array_equal(out_k, all_k[order])
    if oracle is not None:
        wk, _ = _oracle_sorted(oracle, all_k, all_r)
        assert np.array_equal(out_k, wk), "keys differ from the oracle's global sort"
    assert np.array_equal(out_r[np.lexsort((out_r, out_k))], all_r[order])
    assert all(r[it][4] for r in res)


@pytest.mark.parametrize("exchange,world", [("pipelined", 2), ("pipelined", 3), ("peer", 2)])
@pytest.mark.parametrize("kind", ["uniform", "low24", "midbits", "sorted", "dup16"])
def test_sharded_sorter_ranks_sharing_one_gpu(gpu, oracle, kind, exchange, world):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_shared_gpu_worker, args=(world, _free_port(), kind, 300_003, exchange, result), nprocs=world, join=True)
    res = [result[r] for r in range(world)]
    for it in range(2):
        _check_global(res, it, oracle)
    # balance: bins are never split, so only uniform keys promise near-equal shares
    if kind == "uniform":
        sizes = [r[0][0].size for r in res]
        assert max(sizes) - min(sizes) < 0.05 * sum(sizes)


@pytest.mark.parametrize("exchange", ["pipelined", "peer", "nccl"])
@pytest.mark.parametrize("kind", ["uniform", "sorted", "low24", "midbits"])
def test_sharded_sorter_two_gpus_nccl(gpu, kind, exchange):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("box has one GPU; the 2-rank path is covered on CPU by test_distributed_gloo.py")
    world = 2
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, _free_port(), kind, 1_000_003, exchange, result), nprocs=world, join=True)
    res = [result[r] for r in range(world)]
    all_k = np.concatenate([r[2] for r in res])
    all_r = np.concatenate([r[3] for r in res])
    out_k = np.concatenate([r[0] for r in res])
    out_r = np.concatenate([r[1] for r in res])
    order = np.lexsort((all_r, all_k))
    assert np.array_equal(out_k, all_k[order])
    assert np.array_equal(out_r[np.lexsort((out_r, out_k))], all_r[order])
    assert all(r[4] for r in res)
