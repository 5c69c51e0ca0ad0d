"""One C4 weight VJP call (for ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
x = torch.rand(8, 64, 16, 112, 112, device='cuda').bfloat16()
v = torch.rand(8, 64, 16, 112, 112, device='cuda').bfloat16()
for _ in range(2):
    evaluator.conv_weight_vjp_raw(x, v, (3, 3, 3), 1, (1, 1, 1), 1, 1)
torch.cuda.synchronize()
