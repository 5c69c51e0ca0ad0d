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
low.  KFAC-reduce factors need no all-reduce at all: they are Gram matrices of the window sums, so
ranks all-gather those (``FactorSweep``) and each forms the full-batch factor.  ``backend="gloo"`` with CPU tensors is supported for the host-side logic only (tests inject
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


def global_scales(global_batch: int, output_total: int, dtype: torch.dtype = torch.float32) -> Tuple[float, float]:
    """``(1/B, 1/(B * prod(O)^2))`` rounded like the reference's ``Tensor([..]).to(x.dtype)``
    (``convNd_kfc.py:105``, ``convNd_kfac_reduce.py:106-108``): through fp32, then through ``dtype``."""
    kfc = float(torch.tensor([1.0 / global_batch]).to(dtype))
    red = float(torch.tensor([1.0 / (global_batch * output_total**2)]).to(dtype))
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
        scale_kfc, scale_red = global_scales(global_batch, o_total, x.dtype)
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


def _align(n: int, to: int = 64) -> int:
    return (n + to - 1) // to * to


class FactorSweep:
    """Repeated factor sweeps over fixed layer inputs: CUDA graphs for the launch-bound part, and an exchange
    step that moves as few bytes as the mathematics allows.

    A ``FactorSweep`` is built once for a fixed list of input tensors (their storage is read at every
    ``run()``, like the static inputs of any CUDA-graph training loop):

    * **KFC** factors are sums over samples of ``A x A`` matrices: every factor is written in place into a
      256-byte aligned slice of a flat bucket (no ``cat`` / copy-back), the calls that fill one bucket are
      one CUDA graph, and ``run()`` issues the asynchronous SUM all-reduce of bucket ``b`` right after
      replaying graph ``b``, so the collective of one bucket overlaps the compute of the next.
    * **KFAC-reduce** factors are Gram matrices of the window sums ``u`` (``B x A`` numbers per layer against
      ``A x A`` for the factor).  Ranks exchange ``u`` with ONE all-gather for all layers (6 MB instead of
      482 MB per sweep of ResNet-50) and every rank forms the full-batch factor from the gathered sums
      (``evaluator.kfac_reduce_from_sums``): bit-identical on all ranks and to a single-process evaluation
      of the concatenated batch.  Stage 1 of all layers is one graph at the start of ``run()``; stage 2 is
      one graph at the end, where it covers the tail of the KFC all-reduces.  Layers without a tensor-core
      route for stage 2 (``A < 256``, exact fp32 / fp64 math) use the fused call and the bucket all-reduce.

    Buckets are per dtype; factors are summed in the dtype of their layer input (like the reference, whose
    einsum emits ``x.dtype``): reduce fp32 inputs if the last bits of a 16-bit sum across ranks matter.
    """

    def __init__(self, x_shards: Sequence[Tensor], layers: Sequence[LayerSpec], global_batch: int,
                 kinds: Sequence[str] = ("kfc", "kfac_reduce"), group=None, simplify: bool = True,
                 bucket_bytes: int = 64 << 20, use_graphs: bool = True, two_stage: Optional[bool] = None):
        from einconv_b200 import _native as nat
        from einconv_b200 import evaluator
        from einconv_b200.utils import _tuple, resolve_geometry

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.kinds = tuple(kinds)
        device = x_shards[0].device
        on_gpu = device.type == "cuda"
        if two_stage is None:
            two_stage = on_gpu
        # ---- work items ---------------------------------------------------------------------------
        fused = []     # (kind, x, spec, scale, numel, shape, dtype): fused call into an all-reduce bucket
        staged = []    # (x, spec, scale, plan, A, groups): KFAC-reduce through window sums + all-gather
        order = []     # per (layer, kind): ("fused", index) | ("staged", index) to restore the layer order
        for x, spec in zip(x_shards, layers):
            N = x.dim() - 2
            _, _, _, _, _, out = resolve_geometry(tuple(x.shape[2:]), spec.kernel_size, spec.stride,
                                                  spec.padding, spec.dilation)
            o_total = k_total = 1
            for o in out:
                o_total *= o
            for k in _tuple(spec.kernel_size, N):
                k_total *= k
            A = x.shape[1] // spec.groups * k_total
            scale_kfc, scale_red = global_scales(global_batch, o_total, x.dtype)
            for kind in self.kinds:
                plan = None
                if kind == "kfac_reduce" and two_stage and x.shape[0] > 0:
                    plan = evaluator.call_plan(nat.OP_KFAC_REDUCE, x.shape[0], spec.groups, x.shape[1] // spec.groups,
                                               0, x.shape[2:], spec.kernel_size, spec.stride, spec.padding,
                                               spec.dilation, x.dtype, simplify)
                    if not evaluator.kfac_reduce_from_sums_supported(plan, self.world):
                        plan = None
                if plan is not None:
                    order.append(("staged", len(staged), kind))
                    staged.append((x, spec, scale_red, plan, A, spec.groups))
                else:
                    order.append(("fused", len(fused), kind))
                    fused.append((kind, x, spec, scale_kfc if kind == "kfc" else scale_red,
                                  spec.groups * A * A, (spec.groups, A, A), x.dtype))
        # ---- flat all-reduce buckets (per dtype, 256-byte aligned slices) ----------------------------
        self.buckets: List[Tensor] = []
        self._calls: List[List[Callable]] = []
        fused_views: List[Optional[Tensor]] = [None] * len(fused)
        fns = {"kfc": evaluator.kfc, "kfac_reduce": evaluator.kfac_reduce}

        def close(seg, dtype):
            if not seg:
                return
            total = sum(_align(it[1][4]) for it in seg)
            bucket = torch.empty(total, dtype=dtype, device=device)
            calls, offset = [], 0
            for idx, (kind, x, spec, scale, numel, shape, _) in seg:
                view = bucket[offset: offset + numel].view(shape)
                offset += _align(numel)
                fused_views[idx] = view

                def call(fn=fns[kind], x=x, spec=spec, scale=scale, view=view):
                    fn(x, spec.kernel_size, stride=spec.stride, padding=spec.padding, dilation=spec.dilation,
                       groups=spec.groups, simplify=simplify, scale=scale, out=view)

                calls.append(call)
            self.buckets.append(bucket)
            self._calls.append(calls)

        for dtype in dict.fromkeys(it[6] for it in fused):
            esz = torch.empty((), dtype=dtype).element_size()
            seg, seg_bytes = [], 0
            # largest factors first: their all-reduces start early and the LAST bucket, whose collective
            # nothing but stage 2 can hide, holds the small ones
            for idx, it in sorted(enumerate(fused), key=lambda pair: -pair[1][4]):
                if it[6] != dtype:
                    continue
                seg.append((idx, it))
                seg_bytes += _align(it[4]) * esz
                if seg_bytes >= bucket_bytes:
                    close(seg, dtype)
                    seg, seg_bytes = [], 0
            close(seg, dtype)
        # ---- window sums of the staged layers: one flat block per rank, all-gathered ------------------
        # Layers whose factors have the same size are stacked: their window sums are the row blocks of ONE
        # [L * A, B] matrix and stage 2 is ONE grouped launch (groups = L) writing [L, A, A] — 8 launches
        # instead of 53 for ResNet-50, each of which fills the machine (a single 1-K-block launch costs
        # ~20 us whatever its size).
        self._stage1: List[Callable] = []
        self._stage2: List[Callable] = []
        staged_out: List[Optional[Tensor]] = [None] * len(staged)
        self.u_all = self.u_local = None
        if staged:
            families: Dict[tuple, List[int]] = {}
            for idx, (x, spec, scale, plan, A, groups) in enumerate(staged):
                key = (A, x.shape[0], x.dtype, scale) if groups == 1 else ("own", idx)
                families.setdefault(key, []).append(idx)
            layout, total = [], 0  # (member indices, offset of the family's block)
            for key, members in families.items():
                x0, _, _, _, A, groups = staged[members[0]]
                layout.append((members, total))
                total += _align(len(members) * groups * A * x0.shape[0])
            self.u_all = torch.zeros((self.world, total), dtype=torch.float32, device=device)
            rank = dist.get_rank(group) if self.world > 1 else 0
            self.u_local = self.u_all[rank]
            for members, off in layout:
                x0, spec0, scale, plan0, A, groups = staged[members[0]]
                b_local, L = x0.shape[0], len(members)
                rows = groups * A
                for pos, idx in enumerate(members):
                    x, spec = staged[idx][0], staged[idx][1]
                    lo = off + pos * rows * b_local
                    u_view = self.u_local[lo: lo + rows * b_local].view(rows, b_local)

                    def s1(x=x, spec=spec, u_view=u_view):
                        evaluator.kfac_reduce_sums(x, spec.kernel_size, stride=spec.stride, padding=spec.padding,
                                                   dilation=spec.dilation, groups=spec.groups, out=u_view)

                    self._stage1.append(s1)
                if L == 1:
                    plan = plan0
                else:
                    # the stacked problem: L "groups" of A channels each over 1-d images of b_local positions
                    plan = evaluator.call_plan(nat.OP_KFAC_REDUCE, b_local, L, A, 0, (1,), 1, 1, 0, 1, x0.dtype, simplify)
                    if not evaluator.kfac_reduce_from_sums_supported(plan, self.world):
                        raise RuntimeError("stacked KFAC-reduce stage 2 unexpectedly unsupported")
                out = torch.empty((L * groups, A, A), dtype=x0.dtype, device=device)
                for pos, idx in enumerate(members):
                    staged_out[idx] = out[pos * groups: (pos + 1) * groups]

                def s2(plan=plan, scale=scale, out=out, off=off, total=total):
                    evaluator.kfac_reduce_from_sums(self.u_all.view(-1)[off:], plan, self.world, total, scale, out)

                self._stage2.append(s2)
        # ---- results in layer order ---------------------------------------------------------------------
        self.results: Dict[str, List[Tensor]] = {k: [] for k in self.kinds}
        for where, idx, kind in order:
            self.results[kind].append(fused_views[idx] if where == "fused" else staged_out[idx])
        # ---- warm up eagerly (lazy module loading, cudaFuncSetAttribute), then capture the graphs ------
        self._graphs = None
        self._graph_s1 = self._graph_s2 = None
        self.launches_per_run = 0  # kernels of this library inside one run() (graph replays hide them)
        if use_graphs and on_gpu:
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            n0 = nat.launch_count()
            with torch.cuda.stream(side):
                for call in self._stage1:
                    call()
                for calls in self._calls:
                    for call in calls:
                        call()
                for call in self._stage2:
                    call()
            self.launches_per_run = nat.launch_count() - n0
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            self._graphs = []
            pool = None

            def capture(calls):
                nonlocal pool
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, pool=pool):
                    for call in calls:
                        call()
                pool = graph.pool()
                return graph

            if self._stage1:
                self._graph_s1 = capture(self._stage1)
            for calls in self._calls:
                self._graphs.append(capture(calls))
            if self._stage2:
                self._graph_s2 = capture(self._stage2)

    def _play(self, graph, calls):
        if graph is not None:
            graph.replay()
        else:
            for call in calls:
                call()

    def run(self) -> Dict[str, List[Tensor]]:
        """One sweep: returns ``{kind: [factor per layer]}`` (views into the buckets / fixed tensors,
        overwritten by the next ``run()``), identical on every rank."""
        works = []
        gather = None
        if self._stage1:
            self._play(self._graph_s1, self._stage1)
            if self.world > 1:
                # in place: rank r's block is row r of u_all
                gather = dist.all_gather_into_tensor(self.u_all.view(-1), self.u_local, group=self.group, async_op=True)
        for b, bucket in enumerate(self.buckets):
            self._play(self._graphs[b] if self._graphs is not None else None, self._calls[b])
            if self.world > 1:
                works.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        if self._stage2:
            if gather is not None:
                gather.wait()
            self._play(self._graph_s2, self._stage2)
        for work in works:
            work.wait()
        return self.results
