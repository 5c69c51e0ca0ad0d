"""CPU, world_size 2 over gloo: the batch-sharding host logic of the Kronecker-factor path.

The product compute path is CUDA-only, so the per-rank factor computation is injected (the C oracle);
what is under test is the sharding arithmetic of einconv_b200.distributed: contiguous shard bounds,
GLOBAL scales on every rank (convNd_kfc.py:105, convNd_kfac_reduce.py:106-108), bucketed asynchronous
SUM all-reduce, and that every rank ends up with the single-process factors of the full batch.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from einconv_b200 import distributed as D
from oracle import direct

LAYERS = [
    (D.LayerSpec(kernel_size=3, stride=1, padding=1), (6, 4, 9, 9)),
    (D.LayerSpec(kernel_size=(3, 2), stride=2, padding=(1, 0), groups=2), (6, 4, 10, 8)),
    (D.LayerSpec(kernel_size=1), (6, 5, 7, 7)),
]


def _oracle_fn(kind):
    fn = direct.kfc if kind == "kfc" else direct.kfac_reduce

    def compute(x, kernel_size, stride=1, padding=0, dilation=1, groups=1, simplify=True, scale=None):
        out = fn(x.numpy(), kernel_size, stride=stride, padding=padding, dilation=dilation,
                 groups=groups, scale=scale)
        return torch.from_numpy(out).float()

    return compute


def _inputs():
    torch.manual_seed(0)
    return [torch.rand(*shape) for _, shape in LAYERS]


def _worker(rank, world, port, bucket_bytes, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xs = _inputs()
        B = xs[0].shape[0]
        lo, hi = D.shard_bounds(B, world, rank)
        out = D.sharded_factors(
            [x[lo:hi] for x in xs], [spec for spec, _ in LAYERS], global_batch=B,
            compute={"kfc": _oracle_fn("kfc"), "kfac_reduce": _oracle_fn("kfac_reduce")},
            bucket_bytes=bucket_bytes,
        )
        results[rank] = {k: [t.numpy() for t in v] for k, v in out.items()}
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_the_batch():
    for batch, world in [(256, 8), (10, 4), (3, 8), (7, 2)]:
        spans = [D.shard_bounds(batch, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("bucket_bytes", [1 << 30, 1])  # one flat bucket / one collective per factor
def test_sharded_factors_match_single_process(bucket_bytes):
    world = 2
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), bucket_bytes, results), nprocs=world, join=True)
    xs = _inputs()
    for kind in ("kfc", "kfac_reduce"):
        fn = direct.kfc if kind == "kfc" else direct.kfac_reduce
        for n, ((spec, _), x) in enumerate(zip(LAYERS, xs)):
            want = fn(x.numpy(), spec.kernel_size, stride=spec.stride, padding=spec.padding,
                      dilation=spec.dilation, groups=spec.groups)
            for rank in range(world):
                got = results[rank][kind][n]
                np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-9, err_msg=f"{kind} layer {n} rank {rank}")
            np.testing.assert_array_equal(results[0][kind][n], results[1][kind][n])
