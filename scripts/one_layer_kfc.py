"""KFC of one ResNet-50 layer shape each, batch 256, one call per layer (for ncu launch lists).
usage: python scripts/one_layer_kfc.py"""
import sys

import torch

sys.path.insert(0, ".")
from bench import RESNET50_LAYERS  # noqa: E402
from einconv_b200 import evaluator  # noqa: E402

evaluator.set_fp32_math("tf32")
for c, h, k, s, p, n in RESNET50_LAYERS:
    x = torch.rand(256, c, h, h, device="cuda")
    for _ in range(2):
        (evaluator.kfac_reduce if len(sys.argv) > 1 and sys.argv[1] == "kfacr" else evaluator.kfc)(x, k, stride=s, padding=p)
    torch.cuda.synchronize()
    del x
print("ok")
