"""Run under torchrun (>= 2 ranks, one GPU each): the exchanged factors of a batch-sharded FactorSweep against
a single-GPU evaluation of the concatenated batch (missing #6).  Exit code 0 = parity."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from einconv_b200 import distributed as D  # noqa: E402
from einconv_b200 import evaluator  # noqa: E402

LAYERS = [(3, 32, 7, 2, 3), (64, 28, 3, 1, 1), (48, 14, 1, 1, 0), (256, 14, 3, 1, 1), (256, 14, 1, 2, 0),
          (512, 7, 3, 1, 1), (300, 7, 1, 1, 0)]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    global_batch = 8 * world
    failures = []
    for mode, tol in (("tf32", 1e-3), ("fp32", 2e-5)):
        evaluator.set_fp32_math(mode)
        full, shards, specs = [], [], []
        for n, (c, h, k, s, p) in enumerate(LAYERS):
            gen = torch.Generator().manual_seed(n)
            x = torch.rand(global_batch, c, h, h, generator=gen).to(device)
            lo, hi = D.shard_bounds(global_batch, world, rank)
            full.append(x)
            shards.append(x[lo:hi].contiguous())
            specs.append(D.LayerSpec(k, stride=s, padding=p))
        for use_graphs in (True, False):
            sweep = D.FactorSweep(shards, specs, global_batch, bucket_bytes=4 << 20, use_graphs=use_graphs)
            for trial in range(2):
                out = sweep.run()
                torch.cuda.synchronize()
                for kind, fn in (("kfc", evaluator.kfc), ("kfac_reduce", evaluator.kfac_reduce)):
                    for n, ((c, h, k, s, p), x) in enumerate(zip(LAYERS, full)):
                        want = fn(x, k, stride=s, padding=p).double()
                        got = out[kind][n].double()
                        err = float((got - want).abs().max() / want.abs().max())
                        if not err <= tol:
                            failures.append((mode, use_graphs, kind, LAYERS[n], err))
                        # every rank holds the same bits
                        ref = out[kind][n].clone()
                        dist.broadcast(ref, src=0)
                        if not torch.equal(ref, out[kind][n]):
                            failures.append((mode, use_graphs, kind, LAYERS[n], "ranks differ"))
        staged = len(sweep._stage1)
        if rank == 0:
            print(f"mode {mode}: {staged} of {len(LAYERS)} KFAC-reduce factors through the all-gather of window sums")
    evaluator.set_fp32_math("auto")
    dist.barrier()
    if rank == 0:
        print("FAILURES" if failures else "nccl factor parity ok", failures)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
