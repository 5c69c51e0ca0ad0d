"""Host -> GPU -> host unfold of C2 through unfoldNd_host for several chunk sizes (wall clock)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from einconv_b200.functionals import unfoldNd_host  # noqa: E402

x = torch.rand(16, 32, 16, 56, 56).pin_memory()
kw = dict(kernel_size=3, dilation=2, padding=0, stride=2)
out = unfoldNd_host(x, **kw)
work = 249.97e6
for chunk in (1, 2, 4, 8, 16):
    for _ in range(2):
        unfoldNd_host(x, out=out, chunk=chunk, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        unfoldNd_host(x, out=out, chunk=chunk, **kw)
        torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / n
    print(f"chunk {chunk:2d}: {ms:.2f} ms  {work / ms / 1e6:.1f} GB/s")
# plain copies for reference
d = torch.empty(16, 864, 4056, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    out.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / 5
print(f"plain D2H of the 224 MB result: {ms:.2f} ms  {224.28 / ms:.1f} GB/s")
