"""Bring-up of the TMA-fed forward / input-VJP kernel (csrc/conv_tma.cu) against torch on the GPU."""
import os
import sys
import torch
import torch.nn.functional as F

import einconv_b200 as ec
from einconv_b200 import evaluator as ev

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def check(name, B, Cin, Cout, size, k, pad, dil=1, groups=1, dtype=torch.float32, bias=True):
    torch.manual_seed(0)
    N = len(size)
    x = torch.randn(B, Cin, *size, device=dev)
    w = torch.randn(Cout, Cin // groups, *([k] * N), device=dev) / (Cin * k**N) ** 0.5
    b = torch.randn(Cout, device=dev) if bias else None
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[N]
    xq, wq = x.to(dtype), w.to(dtype)
    bq = b.to(dtype) if bias else None
    ref = conv(xq.double(), wq.double(), bq.double() if bias else None, padding=pad, dilation=dil, groups=groups)
    ev.set_fp32_math("tf32" if dtype == torch.float32 else "auto")
    name_k = "?"
    got = ec.functionals.convNd(xq, wq, bq, padding=pad, dilation=dil, groups=groups)
    torch.cuda.synchronize()
    err = (got.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"{name}: max|err|={err:.3e} rel={err/scale:.3e}")
    # input VJP
    v = torch.randn_like(ref).to(dtype)
    xg = xq.double().requires_grad_(True)
    out = conv(xg, wq.double(), None, padding=pad, dilation=dil, groups=groups)
    (gref,) = torch.autograd.grad(out, xg, v.double())
    gx = ev.conv_input_vjp_raw(wq, v, tuple(size), (1,) * N, (pad,) * N if isinstance(pad, int) else pad, (dil,) * N, groups)
    torch.cuda.synchronize()
    err = (gx.double() - gref).abs().max().item()
    print(f"{name}: input VJP max|err|={err:.3e} rel={err/gref.abs().max().item():.3e}")


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("small", "all"):
    check("tf32 2d 8ch", 2, 32, 32, (12, 12), 3, 1)
    check("bf16 2d", 2, 64, 64, (12, 12), 3, 1, dtype=torch.bfloat16)
if which in ("c1", "all"):
    check("C1 tf32", 32, 64, 64, (56, 56), 3, 1)
    check("C1 bf16", 32, 64, 64, (56, 56), 3, 1, dtype=torch.bfloat16)
if which in ("more", "all"):
    check("tf32 1d", 3, 16, 24, (50,), 5, 2)
    check("tf32 3d dil2", 2, 16, 40, (9, 10, 11), 3, 2, dil=2)
    check("tf32 groups", 2, 64, 32, (14, 14), 3, 1, groups=2)
    check("tf32 c3", 2, 3, 20, (20, 20), 7, 3)
    check("f16 cin300", 2, 300, 520, (7, 7), 3, 1, dtype=torch.float16)
