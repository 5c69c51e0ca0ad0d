"""Where does the 1x1 KFC of a 64-channel layer lose its bandwidth?  Vary dtype / batch / kernel variant."""
import os, sys
import torch
sys.path.insert(0, ".")
from einconv_b200 import evaluator
evaluator.set_fp32_math("tf32")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n * 1e3
for env in ({}, {"EINCONV_KFC_NO_DEEP": "1"}, {"EINCONV_KFC_TILE128": "1"}, {"EINCONV_KFC_NO_SPLIT": "1"}):
    for k in ("EINCONV_KFC_NO_DEEP", "EINCONV_KFC_TILE128", "EINCONV_KFC_NO_SPLIT"): os.environ.pop(k, None)
    os.environ.update(env)
    evaluator.clear_plan_cache()
    for (B, C, H, dt) in ((256, 64, 56, torch.float32), (256, 64, 56, torch.bfloat16), (1024, 64, 56, torch.float32),
                          (256, 128, 56, torch.float32), (256, 256, 56, torch.float32), (256, 64, 112, torch.float32)):
        x = torch.rand(B, C, H, H, device="cuda").to(dt)
        us = t(lambda: evaluator.kfc(x, 1))
        print(f"{env or 'default'}: B{B} C{C} H{H} {str(dt)[6:]:8s} {us:8.1f} us  {x.numel() * x.element_size() / us / 1e6:6.2f} TB/s", flush=True)
        del x
