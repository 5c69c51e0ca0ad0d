"""KFC of the large-A ResNet-50 layers at the 8-rank shard size (batch 32), for ncu launch lists."""
import sys, torch
sys.path.insert(0, ".")
from einconv_b200 import evaluator
evaluator.set_fp32_math("tf32")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for c, h, k, s, p in [(512, 7, 3, 1, 1), (256, 14, 3, 1, 1), (128, 28, 3, 1, 1), (2048, 7, 1, 1, 0), (64, 56, 3, 1, 1), (1024, 14, 1, 1, 0)]:
    x = torch.rand(B, c, h, h, device="cuda")
    for _ in range(2):
        evaluator.kfc(x, k, stride=s, padding=p, scale=1.0 / 256)
    torch.cuda.synchronize()
print("ok")
