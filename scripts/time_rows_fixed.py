"""Fixed cost of the row kernel: graph of 10 back-to-back calls at small batches (bring-up build with
EINCONV_ROWS_SKIP_PACK=1 launches only the main kernel)."""
import os, sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
from einconv_b200.functionals import convNd
evaluator.set_fp32_math('tf32')
dev = 'cuda'
sp = torch.rand(8192, 8192, device=dev)
def spin():
    for _ in range(200): sp @ sp
    torch.cuda.synchronize()
for B in (1, 4, 5, 6, 11, 16, 32):
    x = torch.rand(B, 64, 56, 56, device=dev); w = torch.rand(64, 64, 3, 3, device=dev) - 0.5
    f = lambda: convNd(x, w, None, padding=1)
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): f()
    spin()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): g.replay()
    e.record(); e.synchronize()
    print(f"batch {B:2d} ({B * 28} tiles): {s.elapsed_time(e) / 200 * 1e3:.2f} us per call")
