"""C4 weight VJP (bf16, [8,64,16,112,112], k3 p1) and a few ResNet-like shapes: ms per call and TFLOP/s.
Routing variants are chosen through the environment (one process per variant):
  EINCONV_WGRAD_NO_FUSED=1   packed channels-last copies + TMA boxes
  EINCONV_WGRAD_NB=2         two taps per unit (N = 128 MMAs)"""
import os, sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
dev = 'cuda'
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n
cases = {
    'C4 (8,64,16,112,112) k3': ((8, 64, 16, 112, 112), 64, (3, 3, 3), (1, 1, 1)),
    '(32,64,56,56) k3 -> 64': ((32, 64, 56, 56), 64, (3, 3), (1, 1)),
    '(32,128,56,56) k3 -> 128': ((32, 128, 56, 56), 128, (3, 3), (1, 1)),
    '(32,256,56,56) k1 -> 64': ((32, 256, 56, 56), 64, (1, 1), (0, 0)),
    '(64,64,32,32) k5 -> 64': ((64, 64, 32, 32), 64, (5, 5), (2, 2)),
}
tag = ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('EINCONV_'))
print('variant:', tag or 'default')
for name, (xs, co, k, p) in cases.items():
    x = torch.rand(*xs, device=dev).bfloat16()
    v = torch.rand(xs[0], co, *xs[2:], device=dev).bfloat16()
    ms = t(lambda: evaluator.conv_weight_vjp_raw(x, v, k, 1, p, 1, 1))
    kt = 1
    for kk in k: kt *= kk
    flops = 2.0 * x.numel() * co * kt
    print(f'  {name}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s')
