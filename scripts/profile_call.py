"""cProfile of the eager public-API call path (host side) for the C3 depthwise call."""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, ".")
from einconv_b200.functionals import convNd
from einconv_b200 import evaluator
xd = torch.rand(32, 256, 28, 28, device="cuda"); wd = torch.rand(256, 1, 3, 3, device="cuda")
f = lambda: convNd(xd, wd, padding=1, groups=256)
for _ in range(100): f()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5000): f()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
# pieces
import timeit
from einconv_b200 import _native as nat
dev = xd.device
def T(name, fn, n=20000):
    t = timeit.timeit(fn, number=n) / n * 1e6
    print(f"{name:40s} {t:6.2f} us")
T("torch.empty out", lambda: torch.empty((32, 256, 28, 28), dtype=torch.float32, device=dev))
T("current_stream().cuda_stream", lambda: torch.cuda.current_stream(dev).cuda_stream)
T("torch.cuda.current_device()", lambda: torch.cuda.current_device())
T("torch._C._cuda_getCurrentRawStream", lambda: torch._C._cuda_getCurrentRawStream(0))
T("allow_tf32 read", lambda: torch.backends.cuda.matmul.allow_tf32)
T("x.contiguous()", lambda: xd.contiguous())
T("x.is_contiguous()", lambda: xd.is_contiguous())
T("x.data_ptr()", lambda: xd.data_ptr())
T("x.shape", lambda: xd.shape)
T("tuple(x.shape[2:])", lambda: tuple(xd.shape[2:]))
T("is_grad_enabled", lambda: torch.is_grad_enabled())
T("x.requires_grad", lambda: xd.requires_grad)
lib = nat.lib()
g = evaluator._geom(32, 256, 1, 1, (28, 28), (3, 3), 1, 1, 1)[0]
T("ctypes selected_kernel (5 args)", lambda: lib.einconv_selected_kernel(1, g, 0, 1, 0))
T("ctypes launch_count (0 args)", lambda: lib.einconv_launch_count())
