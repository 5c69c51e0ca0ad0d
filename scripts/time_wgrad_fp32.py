import sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
evaluator.set_fp32_math('tf32')
x = torch.rand(8, 64, 16, 112, 112, device='cuda'); v = torch.rand(8, 64, 16, 112, 112, device='cuda')
def t(n=5):
    evaluator.conv_weight_vjp_raw(x, v, (3,3,3), 1, 1, 1, 1); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): evaluator.conv_weight_vjp_raw(x, v, (3,3,3), 1, 1, 1, 1)
    b.record(); b.synchronize(); return a.elapsed_time(b)/n
ms = t(); print(f"C4 shape fp32 (TF32 mode): {ms:.3f} ms  {355.14/ms:.0f} TFLOP/s")
