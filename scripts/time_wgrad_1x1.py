"""Weight VJP of 1x1 convolutions: slab route (packed copies) against the TMA reduction kernel that reads the
channels-first tensors directly (EINCONV_NO_WGRAD_SLAB=1)."""
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
evaluator.set_fp32_math('tf32')
tag = ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('EINCONV_'))
print('variant:', tag or 'default')
for xs, co, dt in [((32, 256, 56, 56), 64, torch.bfloat16), ((32, 64, 56, 56), 256, torch.bfloat16), ((32, 512, 28, 28), 128, torch.bfloat16),
                   ((32, 1024, 14, 14), 256, torch.bfloat16), ((32, 256, 56, 56), 64, torch.float32), ((32, 512, 28, 28), 128, torch.float32)]:
    x = torch.rand(*xs, device=dev).to(dt)
    v = torch.rand(xs[0], co, *xs[2:], device=dev).to(dt)
    ms = t(lambda: evaluator.conv_weight_vjp_raw(x, v, (1, 1), 1, 0, 1, 1))
    g = evaluator._geom(xs[0], 1, xs[1], co, tuple(xs[2:]), (1, 1), 1, 0, 1)[0]
    from einconv_b200 import _native as nat
    print(f'  {tuple(xs)} -> {co} {str(dt)[6:]}: {ms*1e3:.1f} us  [{evaluator.selected_kernel(nat.OP_CONV_WEIGHT_VJP, g, dt)}]')
