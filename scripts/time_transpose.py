"""Timing of the transpose-convolution kernels (row f2) on a decoder-style layer:
x [32, 64, 28, 28] -> conv_transpose 4x4 stride 2 padding 1 (associated conv input 56x56)."""
import sys
import torch
sys.path.insert(0, ".")
from einconv_b200 import evaluator

dev = "cuda"
x = torch.rand(32, 64, 28, 28, device=dev)
kw = dict(kernel_size=4, stride=2, padding=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1); _ = flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


u = evaluator.unfold_transpose(x, **kw)
ms = t(lambda: evaluator.unfold_transpose(x, **kw))
out_bytes = u.numel() * 4
print(f"unfold_transpose {tuple(x.shape)} -> {tuple(u.shape)}: {ms*1e3:.1f} us, "
      f"{(out_bytes + x.numel()*4)/ms/1e6:.0f} GB/s (out {out_bytes/1e6:.0f} MB + in {x.numel()*4/1e6:.1f} MB)")
ms = t(lambda: evaluator.kfac_reduce_transpose(x, **kw))
print(f"kfac_reduce_transpose A={64*16}: {ms*1e3:.1f} us ({x.numel()*4/ms/1e6:.0f} GB/s of input)")
