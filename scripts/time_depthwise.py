"""Kernel time of the C3 depthwise forward / input VJP, back to back on rotating buffers (cold L2),
and of the two-kernel step replayed from a CUDA graph.  Usage: python scripts/time_depthwise.py"""
import sys

import torch

sys.path.insert(0, ".")
from einconv_b200 import evaluator
from einconv_b200.functionals import convNd

dev = torch.device("cuda:0")
R = 10
xs = [torch.rand(32, 256, 28, 28, device=dev) for _ in range(R)]
vs = [torch.rand(32, 256, 28, 28, device=dev) for _ in range(R)]
w = torch.rand(256, 1, 3, 3, device=dev)
bytes_one = 2 * 32 * 256 * 28 * 28 * 4


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def graph_of(body):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            body()
    return g


def fwd_all():
    for x in xs:
        convNd(x, w, padding=1, groups=256)


def vjp_all():
    for v in vs:
        evaluator.conv_input_vjp_raw(w, v, (28, 28), 1, 1, 1, 256)


def both_all():
    for x, v in zip(xs, vs):
        convNd(x, w, padding=1, groups=256)
        evaluator.conv_input_vjp_raw(w, v, (28, 28), 1, 1, 1, 256)


for name, body, n_k in (("fwd", fwd_all, R), ("input_vjp", vjp_all, R), ("fwd+vjp", both_all, R)):
    g = graph_of(body)
    ms = timed(g.replay) / n_k
    per = bytes_one * (2 if name == "fwd+vjp" else 1)
    print(f"{name}: {ms * 1e3:.1f} us per {'step' if '+' in name else 'kernel'}, {per / ms / 1e6:.0f} GB/s, "
          f"{per / ms / 1e6 / 6547.2:.2f} of HBM")
