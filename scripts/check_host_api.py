import sys
import torch
sys.path.insert(0, ".")
from einconv_b200.functionals import unfoldNd, unfoldNd_host
x = torch.rand(5, 3, 9, 10).pin_memory()
o = unfoldNd_host(x, 3, padding=1, stride=2, chunk=2)
ref = unfoldNd(x.cuda(), 3, padding=1, stride=2).cpu()
print("host api equal:", torch.equal(o, ref), tuple(o.shape))
