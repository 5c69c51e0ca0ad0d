"""One C1 forward call in TF32 mode (for ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
from einconv_b200.functionals import convNd
evaluator.set_fp32_math('tf32')
x = torch.rand(32, 64, 56, 56, device='cuda')
w = torch.rand(64, 64, 3, 3, device='cuda') - 0.5
for _ in range(3):
    convNd(x, w, None, padding=1)
torch.cuda.synchronize()
