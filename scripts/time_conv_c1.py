"""C1 forward (fp32, TF32 mode, [32,64,56,56] k3 p1) and a few other forward shapes: us per call, TFLOP/s.
Routing through the environment, one process per variant: EINCONV_CTMA_NO_FUSED=1 = packed copy + TMA."""
import os, sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
from einconv_b200.functionals import convNd
dev = 'cuda'
_spin = None
def spin(ms=300):
    """sustained load first: a cold GPU sits at its idle clock (~1.2 GHz) for the first milliseconds of work, and
    a 20-call measurement is over before it ramps (seen: 22.8 us per kernel cold, 13 us warm)"""
    global _spin
    if _spin is None: _spin = torch.rand(8192, 8192, device=dev)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    while True:
        for _ in range(10): _spin @ _spin
        e.record(); e.synchronize()
        if s.elapsed_time(e) > ms: return
def t(fn, n=200):
    spin()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n
def tg(fn, n=200):
    """graph replay: kernels back to back without the host"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return t(g.replay, n)
cases = [
    ('C1 fp32/tf32 (32,64,56,56)->64 k3', (32, 64, 56, 56), 64, 3, 1, torch.float32),
    ('bf16 (32,64,56,56)->64 k3', (32, 64, 56, 56), 64, 3, 1, torch.bfloat16),
    ('fp32/tf32 (32,128,28,28)->128 k3', (32, 128, 28, 28), 128, 3, 1, torch.float32),
    ('bf16 (32,256,56,56)->256 k3', (32, 256, 56, 56), 256, 3, 1, torch.bfloat16),
    ('fp32/tf32 (32,256,56,56)->64 k1', (32, 256, 56, 56), 64, 1, 0, torch.float32),
]
evaluator.set_fp32_math('tf32')
tag = ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('EINCONV_'))
print('variant:', tag or 'default')
for name, xs, co, k, p, dt in cases:
    x = torch.rand(*xs, device=dev).to(dt)
    w = (torch.rand(co, xs[1], k, k, device=dev) - 0.5).to(dt)
    f = lambda: convNd(x, w, None, padding=p)
    ms, msg = t(f), tg(f)
    flops = 2.0 * xs[0] * co * xs[1] * k * k * xs[2] * xs[3]
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=p)
    err = (f().double() - ref).abs().max().item() / ref.abs().max().item()
    print(f'  {name}: eager {ms*1e3:.1f} us, graph {msg*1e3:.1f} us = {flops/msg/1e9:.1f} TFLOP/s, err/max {err:.2e}')
