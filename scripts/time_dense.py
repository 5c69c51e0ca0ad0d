"""1x1 convolutions (forward + input VJP): the GEMM on the channels-first tensors (conv_dense.cu) against the
packed-copy route (EINCONV_NO_CONV_DENSE=1).  us per call in a CUDA graph, TFLOP/s, GB/s of x + y."""
import os, sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
from einconv_b200.functionals import convNd
dev = 'cuda'
def tg(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): g.replay()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n
cases = [
    ((32, 256, 56, 56), 64, torch.bfloat16), ((32, 64, 56, 56), 256, torch.bfloat16),
    ((32, 512, 28, 28), 128, torch.bfloat16), ((32, 1024, 28, 28), 256, torch.bfloat16),
    ((32, 256, 56, 56), 64, torch.float32), ((32, 64, 56, 56), 256, torch.float32),
    ((32, 512, 28, 28), 128, torch.float32),
]
evaluator.set_fp32_math('tf32')
tag = ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('EINCONV_'))
print('variant:', tag or 'default')
for xs, co, dt in cases:
    x = torch.rand(*xs, device=dev).to(dt)
    w = (torch.rand(co, xs[1], 1, 1, device=dev) - 0.5).to(dt)
    f = lambda: convNd(x, w, None)
    y = f()
    v = torch.rand_like(y)
    fi = lambda: evaluator.conv_input_vjp_raw(w, v, tuple(xs[2:]), 1, 0, 1, 1)
    ms, mi = tg(f), tg(fi)
    flops = 2.0 * xs[0] * co * xs[1] * xs[2] * xs[3]
    bytes_ = (x.numel() + y.numel()) * x.element_size()
    print(f'  {tuple(xs)} -> {co} {str(dt)[6:]}: fwd {ms*1e3:.1f} us ({flops/ms/1e9:.0f} TFLOP/s, {bytes_/ms/1e6:.0f} GB/s), '
          f'input VJP {mi*1e3:.1f} us')
