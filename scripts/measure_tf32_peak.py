"""Dense TF32 and bf16 matmul throughput of this box (cuBLAS through torch), measured like
MEASURED_PEAKS.json's bf16 figure: square GEMMs, CUDA events, best of several repetitions."""
import torch

dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = True
for n in (4096, 8192):
    for dt, name in ((torch.float32, "tf32"), (torch.bfloat16, "bf16")):
        a = torch.randn(n, n, device=dev, dtype=dt)
        b = torch.randn(n, n, device=dev, dtype=dt)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                a @ b
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        print(f"{name} {n}^3: {best*1e3:8.1f} us  {2*n**3/best/1e9:8.1f} TFLOP/s")
