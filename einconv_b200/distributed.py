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


class FactorSweep:
    """Repeated factor sweeps over fixed layer inputs, with the launch-bound part in CUDA graphs.

    ``sharded_factors`` pays ~25 us of host work and 4-8 kernel launches per factor; at 8 ranks the
    per-rank batch of a ResNet-50 sweep is so small that this, not the arithmetic, sets the time.
    A ``FactorSweep`` is built once for a fixed list of input tensors (their storage is read at every
    ``run()``, like the static inputs of any CUDA-graph training loop):

    * every factor is written in place into a slice of a flat fp bucket (no ``cat`` / copy-back);
    * the calls that fill one bucket are captured into one CUDA graph;
    * ``run()`` replays graph ``b`` and immediately issues the asynchronous SUM all-reduce of bucket
      ``b`` on NCCL's stream, so the collective of one bucket overlaps the compute of the next.

    The numbers are those of ``sharded_factors`` (same kernels, same order of operations).
    """

    def __init__(self, x_shards: Sequence[Tensor], layers: Sequence[LayerSpec], global_batch: int,
                 kinds: Sequence[str] = ("kfc", "kfac_reduce"), group=None, simplify: bool = True,
                 bucket_bytes: int = 64 << 20, use_graphs: bool = True):
        from einconv_b200 import evaluator
        from einconv_b200.utils import _tuple, resolve_geometry

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.kinds = tuple(kinds)
        device = x_shards[0].device
        dtype = x_shards[0].dtype
        # ---- work items and their output shapes
        items = []  # (kind, x, spec, scale, numel, shape)
        for x, spec in zip(x_shards, layers):
            N = x.dim() - 2
            _, t_k, _, _, _, out = resolve_geometry(tuple(x.shape[2:]), spec.kernel_size, spec.stride,
                                                    spec.padding, spec.dilation)
            o_total = k_total = 1
            for o in out:
                o_total *= o
            for k in _tuple(spec.kernel_size, N):
                k_total *= k
            A = x.shape[1] // spec.groups * k_total
            scale_kfc, scale_red = global_scales(global_batch, o_total)
            for kind in self.kinds:
                items.append((kind, x, spec, scale_kfc if kind == "kfc" else scale_red,
                              spec.groups * A * A, (spec.groups, A, A)))
        # ---- flat buckets
        esz = torch.empty((), dtype=dtype).element_size()
        self.buckets: List[Tensor] = []
        self._segments: List[List[Tuple]] = []
        seg, seg_elems = [], 0
        for it in items:
            seg.append(it)
            seg_elems += it[4]
            if seg_elems * esz >= bucket_bytes:
                self._segments.append(seg)
                self.buckets.append(torch.empty(seg_elems, dtype=dtype, device=device))
                seg, seg_elems = [], 0
        if seg:
            self._segments.append(seg)
            self.buckets.append(torch.empty(seg_elems, dtype=dtype, device=device))
        self.results: Dict[str, List[Tensor]] = {k: [] for k in self.kinds}
        self._calls: List[List[Callable]] = []
        fns = {"kfc": evaluator.kfc, "kfac_reduce": evaluator.kfac_reduce}
        for seg, bucket in zip(self._segments, self.buckets):
            calls, offset = [], 0
            for kind, x, spec, scale, numel, shape in seg:
                view = bucket[offset : offset + numel].view(shape)
                offset += numel
                self.results[kind].append(view)

                def call(fn=fns[kind], x=x, spec=spec, scale=scale, view=view):
                    fn(x, spec.kernel_size, stride=spec.stride, padding=spec.padding, dilation=spec.dilation,
                       groups=spec.groups, simplify=simplify, scale=scale, out=view)

                calls.append(call)
            self._calls.append(calls)
        # ---- warm up eagerly (lazy module loading, cudaFuncSetAttribute), then capture one graph per bucket
        self._graphs = None
        self.launches_per_run = 0  # kernels of this library inside one run() (graph replays hide them)
        if use_graphs and device.type == "cuda":
            from einconv_b200 import _native as nat

            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            n0 = nat.launch_count()
            with torch.cuda.stream(side):
                for calls in self._calls:
                    for call in calls:
                        call()
            self.launches_per_run = nat.launch_count() - n0
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            self._graphs = []
            pool = None
            for calls in self._calls:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, pool=pool):
                    for call in calls:
                        call()
                pool = graph.pool()
                self._graphs.append(graph)

    def run(self) -> Dict[str, List[Tensor]]:
        """One sweep: returns ``{kind: [factor per layer]}`` (views into the buckets, overwritten by
        the next ``run()``), identical on every rank."""
        works = []
        for b, bucket in enumerate(self.buckets):
            if self._graphs is not None:
                self._graphs[b].replay()
            else:
                for call in self._calls[b]:
                    call()
            if self.world > 1:
                works.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for work in works:
            work.wait()
        return self.results
