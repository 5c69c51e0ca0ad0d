"""One call of a BASELINE workload through the public API (for ncu captures).
usage: python scripts/one_call.py {unfold|conv_fwd|depthwise|weight_vjp}"""
import sys

import torch

sys.path.insert(0, ".")
from einconv_b200 import evaluator  # noqa: E402
from einconv_b200.functionals import convNd, unfoldNd  # noqa: E402

what = sys.argv[1]
dev = "cuda"
torch.manual_seed(0)
if what == "unfold":
    x = torch.rand(16, 32, 16, 56, 56, device=dev)
    for _ in range(2):
        y = unfoldNd(x, 3, dilation=2, padding=0, stride=2)
elif what == "conv_fwd":
    evaluator.set_fp32_math("tf32")
    x, w = torch.rand(32, 64, 56, 56, device=dev), torch.rand(64, 64, 3, 3, device=dev)
    for _ in range(2):
        y = convNd(x, w, padding=1)
elif what == "depthwise":
    x, w = torch.rand(32, 256, 28, 28, device=dev), torch.rand(256, 1, 3, 3, device=dev)
    v = torch.rand(32, 256, 28, 28, device=dev)
    for _ in range(2):
        y = convNd(x, w, padding=1, groups=256)
        g = evaluator.conv_input_vjp_raw(w, v, (28, 28), 1, 1, 1, 256)
elif what == "weight_vjp":
    x = torch.rand(8, 64, 16, 112, 112, device=dev).bfloat16()
    v = torch.rand(8, 64, 16, 112, 112, device=dev).bfloat16()
    for _ in range(2):
        y = evaluator.conv_weight_vjp_raw(x, v, (3, 3, 3), 1, 1, 1, 1)
else:
    raise SystemExit(__doc__)
torch.cuda.synchronize()
print(what, "ok", tuple(y.shape))
