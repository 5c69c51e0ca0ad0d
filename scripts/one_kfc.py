"""One KFC call of one layer geometry (for ncu --set full captures).
usage: python scripts/one_kfc.py C H k s p [batch] [kfacr]"""
import sys
import torch
sys.path.insert(0, ".")
from einconv_b200 import evaluator  # noqa: E402
c, h, k, s, p = map(int, sys.argv[1:6])
B = int(sys.argv[6]) if len(sys.argv) > 6 else 256
fn = evaluator.kfac_reduce if len(sys.argv) > 7 else evaluator.kfc
evaluator.set_fp32_math("tf32")
x = torch.rand(B, c, h, h, device="cuda")
for _ in range(2):
    fn(x, k, stride=s, padding=p)
torch.cuda.synchronize()
print("ok")
