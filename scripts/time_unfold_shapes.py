"""unfoldNd throughput on a few shapes beyond the C2 headline (algorithmic bytes = out + in)."""
import sys
import torch
sys.path.insert(0, ".")
from einconv_b200.functionals import unfoldNd

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [
    ("C2 headline", (16, 32, 16, 56, 56), dict(kernel_size=3, dilation=2, stride=2), torch.float32),
    ("C2 bf16", (16, 32, 16, 56, 56), dict(kernel_size=3, dilation=2, stride=2), torch.bfloat16),
    ("resnet 3x3 56", (32, 64, 56, 56), dict(kernel_size=3, padding=1), torch.float32),
    ("resnet 3x3 56 bf16", (32, 64, 56, 56), dict(kernel_size=3, padding=1), torch.bfloat16),
    ("resnet 3x3 14", (32, 256, 14, 14), dict(kernel_size=3, padding=1), torch.float32),
    ("stem 7x7 s2 224", (32, 3, 224, 224), dict(kernel_size=7, stride=2, padding=3), torch.float32),
    ("3x3 224 (tiled)", (8, 16, 224, 224), dict(kernel_size=3, padding=1), torch.float32),
    ("1d k9 4096", (16, 64, 4096), dict(kernel_size=9, padding=4), torch.float32),
    ("3d 3x3x3 s1", (4, 16, 16, 56, 56), dict(kernel_size=3, padding=1), torch.float32),
]
for name, shape, kw, dt in SHAPES:
    x = torch.rand(*shape, device=dev).to(dt)
    u = unfoldNd(x, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        flush.fill_(1); _ = flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); u = unfoldNd(x, **kw); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = (u.numel() + x.numel()) * x.element_size()
    print(f"{name:22s} {str(dt)[6:]:9s} out {u.numel()*x.element_size()/1e6:8.1f} MB  {ms*1e3:8.1f} us  {nbytes/ms/1e6:7.0f} GB/s  {nbytes/ms/1e6/6547.2:5.2f} of HBM")
    del u, x
