"""Host-side cost of one call through the public API (tiny tensors: the GPU work is negligible), and the
eager per-call latency of the C1 / C3 benchmark calls next to their kernel time."""
import sys, time
import torch
sys.path.insert(0, ".")
from einconv_b200.functionals import convNd, unfoldNd
from einconv_b200 import evaluator

x = torch.rand(2, 8, 12, 12, device="cuda")
w = torch.rand(8, 8, 3, 3, device="cuda")
xb, wb = x.bfloat16(), w.bfloat16()
def bench(name, fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    print(f"{name:44s} {(time.perf_counter()-t0)/n*1e6:7.1f} us per call")
bench("unfoldNd", lambda: unfoldNd(x, 3, padding=1))
bench("convNd fp32 (FFMA)", lambda: convNd(x, w, padding=1))
bench("convNd bf16 (conv_tma)", lambda: convNd(xb, wb, padding=1))
bench("kfc fp32", lambda: evaluator.kfc(x, 3, padding=1))
bench("torch conv2d (reference point)", lambda: torch.nn.functional.conv2d(x, w, padding=1))
print(evaluator.plan_cache_info())
# the benchmark calls, eagerly
xd = torch.rand(32, 256, 28, 28, device="cuda"); wd = torch.rand(256, 1, 3, 3, device="cuda")
bench("C3 depthwise fwd (eager, back to back)", lambda: convNd(xd, wd, padding=1, groups=256))
bench("C3 depthwise input VJP (eager)", lambda: evaluator.conv_input_vjp_raw(wd, xd, (28, 28), 1, 1, 1, 256))
evaluator.set_fp32_math("tf32")
x1 = torch.rand(32, 64, 56, 56, device="cuda"); w1 = torch.rand(64, 64, 3, 3, device="cuda")
bench("C1 conv fwd TF32 mode (eager)", lambda: convNd(x1, w1, padding=1))
