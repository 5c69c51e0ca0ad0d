"""Bring-up of the slab weight-VJP kernel (csrc/wgrad_slab.cu) against torch autograd in fp64."""
import sys
import torch
import torch.nn.functional as F

from einconv_b200 import evaluator as ev

dev = torch.device("cuda:0")


def check(name, B, Cin, Cout, size, k, pad, dil=1, groups=1, dtype=torch.bfloat16):
    torch.manual_seed(0)
    N = len(size)
    ks = (k,) * N if isinstance(k, int) else k
    x = torch.randn(B, Cin, *size, device=dev).to(dtype)
    w = torch.zeros(Cout, Cin // groups, *ks, device=dev, dtype=torch.float64, requires_grad=True)
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[N]
    y = conv(x.double(), w, None, padding=pad, dilation=dil, groups=groups)
    v = torch.randn_like(y).to(dtype)
    (ref,) = torch.autograd.grad(y, w, v.double())
    tp = lambda a: (a,) * N if isinstance(a, int) else tuple(a)
    got = ev.conv_weight_vjp_raw(x, v, ks, tp(dil), tp(pad), (1,) * N, groups)
    torch.cuda.synchronize()
    err = (got.double() - ref).abs().max().item()
    print(f"{name}: max|err|={err:.3e} rel={err/ref.abs().max().item():.3e}", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("small", "all"):
    check("2d 64->64", 2, 64, 64, (12, 12), 3, 1)
    check("2d 32->48", 2, 32, 48, (20, 20), 3, 1)
    check("1x1 128->256", 2, 128, 256, (9, 9), 1, 0)
if which in ("more", "all"):
    check("1d", 3, 16, 24, (50,), 5, 2)
    check("3d dil2", 2, 16, 40, (9, 10, 11), 3, 2, dil=2)
    check("groups", 2, 64, 32, (14, 14), 3, 1, groups=2)
    check("f16 cin300", 2, 304, 520, (7, 7), 3, 1, dtype=torch.float16)
    check("C4 reduced", 2, 64, 64, (8, 112, 112), 3, 1)
