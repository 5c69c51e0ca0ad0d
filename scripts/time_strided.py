"""ResNet-50 stride-2 forward convolutions: TMA kernel over phase sub-images vs the software-gather tcgen05 kernel."""
import os, sys
import torch
sys.path.insert(0, ".")
from einconv_b200 import evaluator
from einconv_b200.functionals import convNd
cases = [((32, 3, 224, 224), (64, 3, 7, 7), 2, 3), ((32, 128, 56, 56), (128, 128, 3, 3), 2, 1), ((32, 256, 56, 56), (512, 256, 1, 1), 2, 0),
         ((32, 256, 28, 28), (256, 256, 3, 3), 2, 1), ((32, 512, 28, 28), (1024, 512, 1, 1), 2, 0), ((32, 512, 14, 14), (512, 512, 3, 3), 2, 1)]
for dt, mode in ((torch.bfloat16, "auto"), (torch.float32, "tf32")):
    evaluator.set_fp32_math(mode)
    for xs, ws, s, p in cases:
        x, w = torch.rand(*xs, device="cuda").to(dt), torch.rand(*ws, device="cuda").to(dt)
        line = f"{str(dt)[6:]:9s} {xs} * {ws} s{s}:"
        for env in (None, "EINCONV_CTMA_NO_STRIDE"):
            os.environ.pop("EINCONV_CTMA_NO_STRIDE", None)
            if env: os.environ[env] = "1"
            evaluator.clear_plan_cache()
            for _ in range(3): convNd(x, w, stride=s, padding=p)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20): convNd(x, w, stride=s, padding=p)
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / 20 * 1e3
            o = (xs[2] + 2 * p - ws[2]) // s + 1
            fl = 2 * xs[0] * ws[0] * ws[1] * ws[2] * ws[3] * o * o
            line += f"  {'gather' if env else 'tma-phases'} {us:7.1f} us {fl / us / 1e6:6.1f} TF/s"
        print(line)
os.environ.pop("EINCONV_CTMA_NO_STRIDE", None)
