"""Bring-up: where does the tf32 MN-major dense kernel see element (channel c0, position p0)?  x one-hot,
w = identity, so y should equal x; prints the (channel, position, value) entries of y that are non-zero."""
import sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
from einconv_b200.functionals import convNd
evaluator.set_fp32_math('tf32')
C, HW = 64, 256
w = torch.eye(C, device="cuda").reshape(C, C, 1).contiguous()
for c0, p0 in [(0, 0), (0, 1), (0, 3), (0, 4), (0, 7), (0, 8), (0, 16), (0, 31), (0, 32), (0, 33), (0, 127), (0, 128),
               (1, 0), (2, 0), (3, 0), (4, 0), (7, 0), (8, 0), (9, 5), (16, 0), (33, 70)]:
    x = torch.zeros(1, C, HW, device='cuda')
    x[0, c0, p0] = 1.0
    y = convNd(x, w, None)
    nz = y[0].nonzero().tolist()
    print((c0, p0), '->', [(c, p, round(float(y[0, c, p]), 3)) for c, p in nz][:6])
