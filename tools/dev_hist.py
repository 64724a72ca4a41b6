"""Developer timing of digit_histogram_kernel (the multi-GPU plan step's device part)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import inplacemsdradixsort_b200 as m
n = 1 << 30
lib = m.load_library()
dev = torch.device("cuda", 0)
k = torch.empty(n, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.msb64_b200_fill(k.data_ptr(), None, n, 0, 1, 0, st)
h = torch.zeros(8194, dtype=torch.int64, device=dev)
best = 1e9
for it in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    lib.msb64_b200_digit_histogram(k.data_ptr(), n, 52, 12, 0, h.data_ptr(), h.data_ptr() + 8 * 4096, st)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print(f"digit_histogram 2^30 keys: {best:.3f} ms, {8 * n / best / 1e6:.0f} GB/s, total {int(h[:4096].sum())}")
