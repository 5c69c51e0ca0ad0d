"""The FactorSweep of bench.py as ONE rank of a `world`-rank job sees it (local batch 256 / world), timed with
CUDA events on a single GPU without any collective: the per-rank compute floor of the multi-GPU sweep.
usage: python scripts/one_sweep.py [world] [reps]   (under ncu: use reps = 1)"""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from bench import FactorsC5, flush_l2  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda")
wl = FactorsC5(world, 0)
wl.setup(dev)
buf = torch.zeros(512 << 20, dtype=torch.uint8, device=dev).view(torch.int32)
for _ in range(2):
    wl.step()
torch.cuda.synchronize()
times = []
for _ in range(reps):
    flush_l2(buf)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    wl.step()
    b.record()
    b.synchronize()
    times.append(a.elapsed_time(b))
print(f"world {world}: local batch {wl.hi - wl.lo}, {statistics.mean(times):.3f} ms per sweep (min {min(times):.3f}), "
      f"{wl.sweep.launches_per_run} launches")
