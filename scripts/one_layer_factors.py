"""One KFC + one KFAC-reduce call per distinct ResNet-50 geometry (for ncu launch lists).
usage: python scripts/one_layer_factors.py [batch]"""
import sys
import torch
sys.path.insert(0, ".")
from bench import RESNET50_LAYERS  # noqa: E402
from einconv_b200 import evaluator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
evaluator.set_fp32_math("tf32")
for c, h, k, s, p, n in RESNET50_LAYERS:
    x = torch.rand(B, c, h, h, device="cuda")
    torch.cuda.synchronize()
    evaluator.kfc(x, k, stride=s, padding=p)
    evaluator.kfac_reduce(x, k, stride=s, padding=p)
    torch.cuda.synchronize()
    del x
print("ok")
