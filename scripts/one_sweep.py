import sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
from bench import RESNET50_LAYERS
evaluator.set_fp32_math('tf32')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
xs = [(torch.rand(B, c, h, h, device='cuda'), k, s, p, n) for c, h, k, s, p, n in RESNET50_LAYERS]
for rep in range(2):
    for x, k, s, p, n in xs:
        for _ in range(n):
            evaluator.kfc(x, k, stride=s, padding=p)
            evaluator.kfac_reduce(x, k, stride=s, padding=p)
torch.cuda.synchronize()
