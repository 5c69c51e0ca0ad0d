"""Ceilings for the unfold kernel on this box: pure-write (fill), copy and read bandwidths, timed
with CUDA events, L2 flushed between repetitions."""
import torch

dev = torch.device("cuda:0")
n = 224280576 // 4
out = torch.empty(n, dtype=torch.float32, device=dev)
src = torch.rand(n, dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    ts = []
    for _ in range(reps + 3):
        flush.fill_(1)
        _ = flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]


t = timeit(lambda: out.fill_(1.0))
print(f"fill   224 MB: {t*1e3:.1f} us  {n*4/t/1e6:.0f} GB/s (write only)")
t = timeit(lambda: out.zero_())
print(f"zero   224 MB: {t*1e3:.1f} us  {n*4/t/1e6:.0f} GB/s (memset)")
t = timeit(lambda: out.copy_(src))
print(f"copy   224 MB: {t*1e3:.1f} us  {2*n*4/t/1e6:.0f} GB/s (read+write)")
t = timeit(lambda: src.sum())
print(f"sum    224 MB: {t*1e3:.1f} us  {n*4/t/1e6:.0f} GB/s (read only)")
big = torch.empty(4 * n, dtype=torch.float32, device=dev)
t = timeit(lambda: big.fill_(1.0))
print(f"fill   897 MB: {t*1e3:.1f} us  {4*n*4/t/1e6:.0f} GB/s (write only)")
big2 = torch.empty(4 * n, dtype=torch.float32, device=dev)
t = timeit(lambda: big2.copy_(big))
print(f"copy   897 MB: {t*1e3:.1f} us  {2*4*n*4/t/1e6:.0f} GB/s (read+write)")
