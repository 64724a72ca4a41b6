"""Developer run of the sharded sort's device steps on ONE GPU (world = 1: the bucket pass runs
with 32 sub-ranges and no peers, then the 32 sub-range sorts): timing, and a target for ncu
(-k regex:bucket_route_kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import inplacemsdradixsort_b200 as m
from bench import parse_count
from inplacemsdradixsort_b200.distributed import ShardedSorter

n = parse_count(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
dev = torch.device("cuda", 0)
lib = m.load_library()
k = torch.empty(n, dtype=torch.int64, device=dev)
r = torch.empty(n, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
assert lib.msb64_b200_fill(k.data_ptr(), r.data_ptr(), n, 0, 7, 0, stream) == 0
with ShardedSorter(n, dev, exchange="pipelined") as s:
    for it in range(3):
        ok_, or_, cnt = s.sort(k, r, timed=True)
        print(it, {a: round(b, 3) if isinstance(b, float) else b for a, b in s.last_times.items()})
    import ctypes
    out = (ctypes.c_uint64 * 3)()
    assert lib.msb64_b200_check(ok_.data_ptr(), or_.data_ptr(), cnt, out, stream) == 0
    assert cnt == n and out[0] == 0, "not sorted"
