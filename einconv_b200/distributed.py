"""Batch-sharded Kronecker factors: one process per GPU, SUM all-reduce over NCCL / NVLink.

The reference has no distributed code.  Its factors are batch averages,
``Omega = 1/B sum_n (...)`` (``einconv/expressions/convNd_kfc.py:103-106``,
``convNd_kfac_reduce.py:105-108``), so they shard by the batch dimension: every rank evaluates the
network on its slice ``x[r*B/R : (r+1)*B/R]`` with the GLOBAL scale, and a plain SUM all-reduce of
the fp32 ``[G, A, A]`` partial factors reproduces the single-process result.  Every other path of
the package (unfold, forward, input VJP) is independent per sample and needs no collective.

Layers are processed as a pipeline: the all-reduce of layer ``l`` is issued asynchronously on
NCCL's stream as soon as its factor kernel has been enqueued, while the compute stream goes on
with layer ``l+1``; small factors are coalesced into flat buckets so the collective count stays
low.  ``backend="gloo"`` with CPU tensors is supported for the host-side logic only (tests inject
the compute function); the product compute path is CUDA-only.
"""

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


@dataclass
class LayerSpec:
    """Geometry of one convolution layer whose input-based factors are wanted."""

    kernel_size: object
    stride: object = 1
    padding: object = 0
    dilation: object = 1
    groups: int = 1


def shard_bounds(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous batch slice ``[lo, hi)`` owned by ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_scales(global_batch: int, output_total: int) -> Tuple[float, float]:
    """``(1/B, 1/(B * prod(O)^2))`` rounded through fp32 like the reference's ``Tensor([..])``."""
    kfc = float(torch.tensor([1.0 / global_batch]))
    red = float(torch.tensor([1.0 / (global_batch * output_total**2)]))
    return kfc, red


def _default_compute(kind: str) -> Callable:
    from einconv_b200 import evaluator

    return evaluator.kfc if kind == "kfc" else evaluator.kfac_reduce


class FactorReducer:
    """Accumulates per-layer partial factors into flat buckets and all-reduces them asynchronously."""

    def __init__(self, group=None, bucket_bytes: int = 64 << 20):
        self.group = group
        self.bucket_bytes = bucket_bytes
        self._pending: List[Tensor] = []
        self._pending_bytes = 0
        self._inflight: List[Tuple[object, Tensor, List[Tensor]]] = []

    def add(self, factor: Tensor) -> None:
        self._pending.append(factor)
        self._pending_bytes += factor.numel() * factor.element_size()
        if self._pending_bytes >= self.bucket_bytes:
            self.flush()

    def flush(self) -> None:
        if not self._pending:
            return
        tensors, self._pending, self._pending_bytes = self._pending, [], 0
        if len(tensors) == 1:
            flat = tensors[0].view(-1)
        else:
            flat = torch.cat([t.reshape(-1) for t in tensors])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight.append((work, flat, tensors))

    def wait(self) -> None:
        self.flush()
        for work, flat, tensors in self._inflight:
            work.wait()
            if len(tensors) > 1:
                offset = 0
                for t in tensors:
                    t.copy_(flat[offset : offset + t.numel()].view_as(t))
                    offset += t.numel()
        self._inflight = []


def sharded_factors(
    x_shards: Sequence[Tensor],
    layers: Sequence[LayerSpec],
    global_batch: int,
    kinds: Sequence[str] = ("kfc", "kfac_reduce"),
    group=None,
    compute: Optional[Dict[str, Callable]] = None,
    simplify: bool = True,
    bucket_bytes: int = 64 << 20,
) -> Dict[str, List[Tensor]]:
    """KFC / KFAC-reduce factors of several layers from this rank's batch shards.

    Args:
        x_shards: Per layer, this rank's slice of the layer input ``[B_local, C_in, *spatial]``.
        layers: Per layer geometry.
        global_batch: ``B`` summed over all ranks (the reference's ``batch_size``).
        kinds: Which factor families to compute.
        group: ``torch.distributed`` process group (default group if ``None``).
        compute: Optional ``{kind: fn(x, kernel_size, stride=, padding=, dilation=, groups=,
            simplify=, scale=)}`` override (CPU tests inject the oracle here).
        simplify: Passed to the evaluator (kernel specialisations on/off).
        bucket_bytes: Flat all-reduce bucket size.

    Returns:
        ``{kind: [factor per layer]}``, identical on every rank, equal to the single-process
        factors of the concatenated batch.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    fns = {k: (compute or {}).get(k) or _default_compute(k) for k in kinds}
    reducer = FactorReducer(group, bucket_bytes) if world > 1 else None
    results: Dict[str, List[Tensor]] = {k: [] for k in kinds}

    for x, spec in zip(x_shards, layers):
        from einconv_b200.utils import resolve_geometry

        _, _, _, _, _, out = resolve_geometry(
            tuple(x.shape[2:]), spec.kernel_size, spec.stride, spec.padding, spec.dilation
        )
        o_total = 1
        for o in out:
            o_total *= o
        scale_kfc, scale_red = global_scales(global_batch, o_total)
        for kind in kinds:
            factor = fns[kind](
                x,
                spec.kernel_size,
                stride=spec.stride,
                padding=spec.padding,
                dilation=spec.dilation,
                groups=spec.groups,
                simplify=simplify,
                scale=scale_kfc if kind == "kfc" else scale_red,
            )
            results[kind].append(factor)
            if reducer is not None:
                reducer.add(factor)
    if reducer is not None:
        reducer.wait()
    return results
