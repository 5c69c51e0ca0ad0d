"""fp32 (TF32 mode) forward convolutions on the TMA kernel: native kind::tf32 vs scaled fp16 operands."""
import os
import sys

import torch

sys.path.insert(0, ".")
from einconv_b200 import evaluator  # noqa: E402
from einconv_b200.functionals import convNd  # noqa: E402

evaluator.set_fp32_math("tf32")
cases = [((32, 64, 56, 56), (64, 64, 3, 3)), ((32, 128, 28, 28), (128, 128, 3, 3)), ((32, 256, 28, 28), (256, 256, 3, 3)),
         ((32, 512, 14, 14), (512, 512, 3, 3)), ((32, 256, 56, 56), (64, 256, 1, 1)), ((32, 64, 56, 56), (256, 64, 1, 1))]
for xs, ws in cases:
    x, w = torch.rand(*xs, device="cuda"), torch.rand(*ws, device="cuda")
    pad = ws[2] // 2
    line = f"{xs} * {ws}:"
    for env in ("EINCONV_CTMA_TF32", "EINCONV_CTMA_HALF"):
        os.environ.pop("EINCONV_CTMA_TF32", None)
        os.environ.pop("EINCONV_CTMA_HALF", None)
        os.environ[env] = "1"
        for _ in range(3):
            convNd(x, w, padding=pad)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            convNd(x, w, padding=pad)
        b.record()
        torch.cuda.synchronize()
        line += f"  {env[-4:].lower()} {a.elapsed_time(b) / 20 * 1e3:.1f} us"
    print(line)
