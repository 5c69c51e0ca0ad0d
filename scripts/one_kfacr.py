import sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
evaluator.set_fp32_math('tf32')
c, h, k, s, p = [int(a) for a in sys.argv[1:6]]
x = torch.rand(256, c, h, h, device='cuda')
for _ in range(3):
    evaluator.kfac_reduce(x, k, stride=s, padding=p)
torch.cuda.synchronize()
