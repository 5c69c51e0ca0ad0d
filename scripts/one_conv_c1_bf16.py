"""One C1-shaped bf16 forward call (for traces / ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
from einconv_b200.functionals import convNd
x = torch.rand(32, 64, 56, 56, device='cuda').bfloat16()
w = (torch.rand(64, 64, 3, 3, device='cuda') - 0.5).bfloat16()
for _ in range(3):
    convNd(x, w, None, padding=1)
torch.cuda.synchronize()
