"""Factor sweep (bench.py's C5 workload) under torchrun with different exchange-bucket sizes: ms per sweep, max over
ranks.  NCCL settings come from the environment of the launch (one process group per run).
usage: torchrun --nproc-per-node N scripts/multi_sweep.py [bucket_mb ...]"""
import os, statistics, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import bench  # noqa: E402
from einconv_b200 import distributed as D, evaluator  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sizes = [int(a) for a in sys.argv[1:]] or [64]
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NCCL_"))
for mb in sizes:
    wl = bench.FactorsC5(world, rank)
    orig = D.FactorSweep.__init__
    def patched(self, *a, **kw):
        kw["bucket_bytes"] = mb << 20
        orig(self, *a, **kw)
    D.FactorSweep.__init__ = patched
    evaluator.set_fp32_math("tf32")
    wl.setup(dev, seed=rank)
    D.FactorSweep.__init__ = orig
    dist.barrier()
    times, _ = bench.time_steps(wl, 10, 3, dev)
    t = torch.tensor([statistics.mean(times)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"[{tag or 'default'}] world {world} bucket {mb} MB ({len(wl.sweep.buckets)} buckets): {t.item():.3f} ms per sweep", flush=True)
    del wl
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
