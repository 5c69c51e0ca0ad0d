"""One 1x1 forward call (64 -> 256 channels, bf16, 56x56) for ncu captures."""
import sys, torch
sys.path.insert(0, '.')
from einconv_b200.functionals import convNd
ci, co = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 256)
x = torch.rand(32, ci, 56, 56, device='cuda').bfloat16()
w = (torch.rand(co, ci, 1, 1, device='cuda') - 0.5).bfloat16()
for _ in range(3):
    convNd(x, w, None)
torch.cuda.synchronize()
