"""C1-shaped input VJP (fp32, TF32 mode): us per call through the row kernel and the packed-copy TMA kernel."""
import os, sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
dev = 'cuda'
def t(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n * 1e3
def tg(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return t(g.replay, n)
evaluator.set_fp32_math('tf32')
for name, xs, co, k, p in [('C1', (32, 64, 56, 56), 64, 3, 1), ('128ch 28x28', (32, 128, 28, 28), 128, 3, 1)]:
    w = torch.rand(co, xs[1], k, k, device=dev) - 0.5
    v = torch.rand(xs[0], co, xs[2], xs[3], device=dev)
    fn = lambda: evaluator.conv_input_vjp(w, v, xs[2:], padding=p)
    flops = 2.0 * xs[0] * xs[1] * co * k * k * xs[2] * xs[3]
    e, g = t(fn), tg(fn)
    print(f"{name} input VJP fp32/tf32: eager {e:.1f} us, graph {g:.1f} us = {flops / g / 1e6:.1f} TFLOP/s")
