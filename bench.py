#!/usr/bin/env python
"""Benchmark of the einconv hot path on B200 (see BASELINE.md for workloads and algorithmic work).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload unfold|conv_fwd|depthwise|weight_vjp|factors]
                    [--impl ours|reference] [--no-suite] [--no-cpu-baseline]

Prints ONE JSON line.  The headline workload is BASELINE.json ``configs[1]`` (C2: unfoldNd 3-d,
N=16 C=32 16x56x56 k=3 stride=2 dilation=2, fp32, bit-exact gather) with metric "unfoldNd HBM GB/s"
= algorithmic bytes (249.97 MB per call: 224.28 MB written + 25.69 MB compulsory read) / time.
A "step" is one call over one synthetic batch.

* ``value``: CUDA events on the launch stream, inputs resident in HBM, L2 flushed between steps.
* ``e2e``: the same call through the public API with HOST (pinned) buffers — the host->device copy
  of the inputs and the device->host copy of the result are inside the timed region.
* ``roofline``: algorithmic bytes / flops over the measured time against MEASURED_PEAKS.json (HBM,
  bf16) and a TF32 cuBLAS peak measured live in this run; ``traffic`` from the committed ncu captures.
* ``cpu_baseline``: the reference's own CPU einsum path (``oracle/_ref`` = the unmodified reference,
  byte-compiled; falls back to the port ``oracle/einsum_port.py`` if it was not built) on the box's
  host cores, on a bounded sample of the same workload.
* ``suite``: the other BASELINE configs (C1, C3, C4, C5), each with the same four measurements.

With ``--gpus N`` (launched by torchrun) every rank runs the workload on its own batch shard
("weak" scaling); the factor workload splits a fixed batch of 256 and adds the exchange step of
SURVEY.md 8e.  Times are the max over ranks.  ``--impl reference`` times the reference's CPU path
on the full C2 batch with every host thread.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

if "reference" in sys.argv:
    # the CPU arm uses every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, and the OpenMP /
    # MKL pools read it when torch is imported
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import torch  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

RESNET50_LAYERS = [
    # (C_in, H_in, k, stride, padding, count)  — SURVEY.md Appendix C
    (3, 224, 7, 2, 3, 1), (64, 56, 1, 1, 0, 5), (64, 56, 3, 1, 1, 3), (256, 56, 1, 1, 0, 3),
    (128, 56, 3, 2, 1, 1), (256, 56, 1, 2, 0, 1), (128, 28, 1, 1, 0, 4), (128, 28, 3, 1, 1, 3),
    (512, 28, 1, 1, 0, 4), (256, 28, 3, 2, 1, 1), (512, 28, 1, 2, 0, 1), (256, 14, 1, 1, 0, 6),
    (256, 14, 3, 1, 1, 5), (1024, 14, 1, 1, 0, 6), (512, 14, 3, 2, 1, 1), (1024, 14, 1, 2, 0, 1),
    (512, 7, 1, 1, 0, 3), (512, 7, 3, 1, 1, 2), (2048, 7, 1, 1, 0, 2),
]


def peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        p = json.loads(path.read_text())
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="MEASURED_PEAKS.json")
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


_TF32_PEAK = {}


def tf32_peak(device):
    """Dense TF32 throughput of THIS box, measured like MEASURED_PEAKS.json's bf16 figure:
    torch.matmul fp32 with allow_tf32, 8192^3, best of 10, CUDA events."""
    key = str(device)
    if key not in _TF32_PEAK:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            n = 8192
            a = torch.randn(n, n, device=device)
            b = torch.randn(n, n, device=device)
            for _ in range(3):
                a @ b
            best = None
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                a @ b
                e1.record()
                e1.synchronize()
                t = e0.elapsed_time(e1)
                best = t if best is None else min(best, t)
            _TF32_PEAK[key] = 2 * n**3 / (best * 1e-3) / 1e12
            del a, b
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        torch.cuda.empty_cache()
    return _TF32_PEAK[key]


# ------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref = unmodified reference; oracle/einsum_port.py = port)
# ------------------------------------------------------------------------------------------------
class ReferenceCPU:
    """Uniform front over the reference's public API for the five workloads."""

    def __init__(self):
        from oracle import ref_build

        self.pkg = ref_build.load()
        self.kind = "reference" if self.pkg is not None else "port"
        if self.pkg is not None:
            import importlib

            self.F = importlib.import_module("einconv.functionals")
            self.E = {n: importlib.import_module(f"einconv.expressions.{n}")
                      for n in ("convNd_input_vjp", "convNd_weight_vjp", "convNd_kfc", "convNd_kfac_reduce")}
        else:
            from oracle import einsum_port

            self.port = einsum_port
        self.source = ("oracle/_ref (unmodified reference, byte-compiled; torch.einsum on the host cores)"
                       if self.pkg is not None else "oracle/einsum_port.py (port of the reference's einsum path)")

    def _expr(self, name, *args, **kw):
        eq, ops, shape = self.E[name].einsum_expression(*args, **kw)
        return torch.einsum(eq, *ops).reshape(shape)

    def unfold(self, x, **kw):
        return self.F.unfoldNd(x, **kw) if self.pkg else self.port.unfoldNd(x, **kw)

    def conv(self, x, w, **kw):
        return self.F.convNd(x, w, **kw) if self.pkg else self.port.convNd(x, w, **kw)

    def input_vjp(self, w, v, input_size, **kw):
        if self.pkg:
            return self._expr("convNd_input_vjp", w, v, input_size, **kw)
        return self.port.evaluate(self.port.input_vjp_expression(w, v, input_size, **kw))

    def weight_vjp(self, x, v, k, **kw):
        if self.pkg:
            return self._expr("convNd_weight_vjp", x, v, k, **kw)
        return self.port.evaluate(self.port.weight_vjp_expression(x, v, k, **kw))

    def kfc(self, x, k, **kw):
        if self.pkg:
            return self._expr("convNd_kfc", x, k, **kw)
        return self.port.evaluate(self.port.kfc_expression(x, k, **kw))

    def kfac_reduce(self, x, k, **kw):
        if self.pkg:
            return self._expr("convNd_kfac_reduce", x, k, **kw)
        return self.port.evaluate(self.port.kfac_reduce_expression(x, k, **kw))


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
class Workload:
    name = ""
    metric = ""
    unit = ""
    bound = "hbm"
    dtype = "f32"
    fp32_math = "fp32"  # how fp32 contractions run: "fp32" (CUDA cores, exact) or "tf32" (tcgen05)
    graph = None
    graph_launches = 0
    ncu_profiles = ()   # committed ncu --set full summaries whose DRAM bytes make up `roofline.traffic`

    def setup(self, device, seed=0):
        raise NotImplementedError

    def step(self):
        raise NotImplementedError

    def work(self):
        """Algorithmic units per step (bytes for GB/s metrics, flops for TFLOP/s, factors)."""
        raise NotImplementedError

    def capture(self, body):
        """Capture `body` (public API calls on fixed buffers) in a CUDA graph: a step of a few ~10 us
        kernels is otherwise bound by the host's launch path, not by the kernels."""
        from einconv_b200 import _native as nat

        if os.environ.get("EINCONV_BENCH_NO_GRAPHS"):
            return
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
            n0 = nat.launch_count()
            body()
            self.graph_launches = nat.launch_count() - n0
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                body()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = g

    # ---- end to end: host buffers in, host buffers out, through the public API -----------------
    def e2e_setup(self, device):
        raise NotImplementedError

    def e2e_step(self, device):
        """One call with pinned host inputs and outputs; returns (h2d bytes, d2h bytes)."""
        raise NotImplementedError

    # ---- the reference on the host cores --------------------------------------------------------
    def cpu_run(self, ref):
        """Returns (seconds, algorithmic units processed, description of the bounded sample)."""
        raise NotImplementedError


def _pinned(t):
    return t.pin_memory() if torch.cuda.is_available() else t


def _nbytes(*tensors):
    return sum(t.numel() * t.element_size() for t in tensors)


class UnfoldC2(Workload):
    name = "C2 unfoldNd 3d N=16 C=32 16x56x56 k=3 stride=2 dilation=2 fp32"
    metric = "unfoldNd HBM GB/s"
    unit = "GB/s"
    bound = "hbm"
    shape = (16, 32, 16, 56, 56)
    kw = dict(kernel_size=3, dilation=2, padding=0, stride=2)
    ncu_profiles = ("profiles/r02_unfold_c2_ncu_full.txt",)

    def setup(self, device, seed=0):
        torch.manual_seed(seed)
        self.x_host = _pinned(torch.rand(*self.shape))
        if device.type == "cuda":
            from einconv_b200.functionals import unfoldNd

            self.fn = unfoldNd
            self.x = self.x_host.to(device)
        self.out = None

    def step(self):
        self.out = self.fn(self.x, **self.kw)

    def work(self):
        out_bytes = 16 * 32 * 27 * (6 * 26 * 26) * 4
        # compulsory input: only even depth planes and even rows are ever read (i = 2o + 2k)
        in_bytes = 16 * 32 * 8 * 28 * 56 * 4
        return out_bytes + in_bytes

    def naive_bytes(self):
        return 16 * 32 * 27 * 4056 * 4 + 16 * 32 * 16 * 56 * 56 * 4

    def kernel_ms(self, device, reps=8):
        """Average duration of the dominant kernel's launches: `reps` calls back to back inside ONE
        CUDA-event pair, each on its own input and output buffers (8 x 103 MB in, 8 x 224 MB out, far
        more than the 126 MB L2), so that every launch finds its input cold without a flush kernel
        between the launches and without the host-side launch gap that a per-step event pair includes."""
        xs = [self.x.clone() for _ in range(reps)]
        outs = [self.fn(x, **self.kw) for x in xs]  # allocations + warm-up
        torch.cuda.synchronize(device)
        best = None
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(reps):
                outs[i] = self.fn(xs[i], **self.kw)
            b.record()
            b.synchronize()
            t = a.elapsed_time(b) / reps
            best = t if best is None else min(best, t)
        return best

    def e2e_setup(self, device):
        self.out_host = torch.empty(16, 864, 4056).pin_memory()

    def e2e_step(self, device):
        # public host-buffer API: batch chunks alternate between two streams so that uploads and the
        # gather overlap the (dominant) download of the previous chunk
        from einconv_b200.functionals import unfoldNd_host

        unfoldNd_host(self.x_host, out=self.out_host, device=device, **self.kw)
        return _nbytes(self.x_host), _nbytes(self.out_host)

    def cpu_run(self, ref):
        t0 = time.perf_counter()
        ref.unfold(self.x_host, **self.kw)
        return time.perf_counter() - t0, self.work(), "the full C2 batch (16/16 samples), simplify=True"


class ConvFwdC1(Workload):
    name = "C1 convNd 2d fwd N=32 C=64->64 56x56 k3 s1 p1 fp32 (step = one CUDA-graph replay of the call)"
    metric = "convNd fwd TFLOP/s"
    unit = "TFLOP/s"
    bound = "tensor"
    dtype = "f32 (tf32 MMA, fp32 accumulate)"
    fp32_math = "tf32"
    ncu_profiles = ("profiles/r02_conv_rows_c1_ncu_full.txt",)  # main kernel; the 147 KB weight pre-pass is not in it

    def setup(self, device, seed=0):
        torch.manual_seed(seed)
        self.x_host = _pinned(torch.rand(32, 64, 56, 56))
        self.w_host = _pinned(torch.rand(64, 64, 3, 3))
        if device.type == "cuda":
            from einconv_b200.functionals import convNd

            self.fn = convNd
            self.x, self.w = self.x_host.to(device), self.w_host.to(device)
            self.capture(self._calls)

    def _calls(self):
        self.out = self.fn(self.x, self.w, padding=1)

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._calls()

    def eager_step(self):
        self._calls()

    def work(self):
        return 2 * 32 * 64 * 64 * 9 * 56 * 56

    def also(self, device, steps=10):
        """The other calls BASELINE's metric names on the C1 shape, each timed like the step (one CUDA-graph replay
        per step, L2 flushed, CUDA events): the input VJP in the same math mode and the bf16 forward
        pass (8d: "also report bf16").  Same flop count as the forward pass."""
        from einconv_b200 import evaluator

        out = {}
        v = torch.rand(32, 64, 56, 56, device=device)
        xb, wb = self.x.bfloat16(), self.w.bfloat16()
        cases = {
            "input_vjp_f32_tf32": lambda: evaluator.conv_input_vjp(self.w, v, (56, 56), padding=1),
            "fwd_bf16": lambda: self.fn(xb, wb, padding=1),
        }
        keep_graph, keep_launches = self.graph, getattr(self, "graph_launches", 0)
        for name, body in cases.items():
            self.graph = None
            self.capture(body)
            replay = self.graph.replay if self.graph is not None else body
            times, _ = time_steps(self, steps, 3, device, step=replay)
            ms = statistics.mean(times)
            out[name] = dict(ms_per_step=ms, value=self.work() / (ms * 1e-3) / 1e12, unit="TFLOP/s",
                             gpu_launches=getattr(self, "graph_launches", None))
        self.graph, self.graph_launches = keep_graph, keep_launches
        return out

    def bytes_algorithmic(self):
        return (2 * 32 * 64 * 56 * 56 + 64 * 64 * 9) * 4

    def e2e_setup(self, device):
        self.y_host = torch.empty(32, 64, 56, 56).pin_memory()

    def e2e_step(self, device):
        x = self.x_host.to(device, non_blocking=True)
        w = self.w_host.to(device, non_blocking=True)
        self.y_host.copy_(self.fn(x, w, padding=1), non_blocking=True)
        return _nbytes(self.x_host, self.w_host), _nbytes(self.y_host)

    def cpu_run(self, ref):
        t0 = time.perf_counter()
        ref.conv(self.x_host, self.w_host, padding=1)
        return time.perf_counter() - t0, self.work(), "the full C1 batch (32/32 samples), fp32, simplify=True"


class DepthwiseC3(Workload):
    name = ("C3 depthwise 2d fwd + input VJP N=32 C=256 28x28 k3 p1 fp32 (step = one CUDA-graph replay of the two "
            "calls, forked streams)")
    metric = "depthwise fwd+input-VJP HBM GB/s"
    unit = "GB/s"
    bound = "hbm"
    ncu_profiles = ("profiles/r02_depthwise_c3_ncu_full.txt",)  # both kernels of the step

    def setup(self, device, seed=0):
        torch.manual_seed(seed)
        self.x_host = _pinned(torch.rand(32, 256, 28, 28))
        self.w_host = _pinned(torch.rand(256, 1, 3, 3))
        self.v_host = _pinned(torch.rand(32, 256, 28, 28))
        if device.type == "cuda":
            from einconv_b200 import evaluator
            from einconv_b200.functionals import convNd

            self.fwd, self.vjp = convNd, evaluator.conv_input_vjp_raw
            self.x, self.w, self.v = self.x_host.to(device), self.w_host.to(device), self.v_host.to(device)
            self.side = None
            self.capture(self._calls)

    def _calls(self):
        # the two directions are independent: the input VJP runs on a forked stream, so that inside the
        # captured graph the two ~10 us kernels overlap instead of queueing behind each other
        cur = torch.cuda.current_stream()
        if self.side is None:
            self.side = torch.cuda.Stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            self.gx = self.vjp(self.w, self.v, (28, 28), 1, 1, 1, 256)
        self.y = self.fwd(self.x, self.w, padding=1, groups=256)
        cur.wait_stream(self.side)

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._calls()

    def eager_step(self):
        # plain user code: two calls on the current stream
        self.gx = self.vjp(self.w, self.v, (28, 28), 1, 1, 1, 256)
        self.y = self.fwd(self.x, self.w, padding=1, groups=256)

    def work(self):
        one = 2 * 32 * 256 * 28 * 28 * 4 + 256 * 9 * 4
        return 2 * one

    def e2e_setup(self, device):
        self.y_host = torch.empty(32, 256, 28, 28).pin_memory()
        self.gx_host = torch.empty(32, 256, 28, 28).pin_memory()

    def e2e_step(self, device):
        x = self.x_host.to(device, non_blocking=True)
        w = self.w_host.to(device, non_blocking=True)
        v = self.v_host.to(device, non_blocking=True)
        self.y_host.copy_(self.fwd(x, w, padding=1, groups=256), non_blocking=True)
        self.gx_host.copy_(self.vjp(w, v, (28, 28), 1, 1, 1, 256), non_blocking=True)
        return _nbytes(self.x_host, self.w_host, self.v_host), _nbytes(self.y_host, self.gx_host)

    def cpu_run(self, ref):
        t0 = time.perf_counter()
        ref.conv(self.x_host, self.w_host, padding=1, groups=256)
        ref.input_vjp(self.w_host, self.v_host, (28, 28), padding=1, groups=256)
        return time.perf_counter() - t0, self.work(), "the full C3 batch (32/32 samples), both directions"


class WeightVjpC4(Workload):
    name = "C4 convNd 3d weight VJP N=8 C=64 16x112x112 k3 p1 bf16"
    metric = "convNd weight-VJP TFLOP/s"
    unit = "TFLOP/s"
    bound = "tensor"
    dtype = "bf16"
    # main kernel + the two channels-last packing launches (x and v) that precede it in every call
    ncu_profiles = ("profiles/r02_wgrad_c4_ncu_full.txt",)  # two packs + slab kernel + finish

    def setup(self, device, seed=0):
        torch.manual_seed(seed)
        self.x_host = _pinned(torch.rand(8, 64, 16, 112, 112).bfloat16())
        self.v_host = _pinned(torch.rand(8, 64, 16, 112, 112).bfloat16())
        if device.type == "cuda":
            from einconv_b200 import evaluator

            self.fn = evaluator.conv_weight_vjp_raw
            self.x, self.v = self.x_host.to(device), self.v_host.to(device)

    def step(self):
        self.out = self.fn(self.x, self.v, 3, 1, 1, 1, 1)

    def work(self):
        return 2 * 8 * 64 * 64 * 27 * 16 * 112 * 112

    def bytes_algorithmic(self):
        return 2 * 8 * 64 * 16 * 112 * 112 * 2 + 64 * 64 * 27 * 2

    def e2e_setup(self, device):
        self.gw_host = torch.empty(64, 64, 3, 3, 3, dtype=torch.bfloat16).pin_memory()

    def e2e_step(self, device):
        x = self.x_host.to(device, non_blocking=True)
        v = self.v_host.to(device, non_blocking=True)
        self.gw_host.copy_(self.fn(x, v, 3, 1, 1, 1, 1), non_blocking=True)
        return _nbytes(self.x_host, self.v_host), _nbytes(self.gw_host)

    def cpu_run(self, ref):
        # one of the 8 samples (the reference's bf16 temporaries for N=8 need > 5 GB)
        x, v = self.x_host[:1], self.v_host[:1]
        t0 = time.perf_counter()
        ref.weight_vjp(x, v, 3, padding=1)
        return time.perf_counter() - t0, self.work() / 8, "1/8 samples of the C4 batch (full 16x112x112 volume), bf16"


class FactorsC5(Workload):
    """KFC + KFAC-reduce over the 53 ResNet-50 conv layers, batch 256 split over the ranks."""

    name = "C5 KFC + KFAC-reduce, 53 ResNet-50 conv layers, batch 256 (sharded by N), fp32"
    metric = "KFC+KFAC-reduce factors/s"
    unit = "factors/s"
    bound = "tensor"
    dtype = "f32 (KFC: tf32 / scaled-fp16 MMA, fp32 accumulate; KFAC-reduce: fp32 sums + tf32 MMA)"
    fp32_math = "tf32"
    # dominant kernel of the sweep: kfc_super_kernel<__half> (a third of the GPU time,
    # profiles/r02_factors_c5_launches_b256_weighted.txt), captured on its most frequent layer
    ncu_profiles = ("profiles/r02_kfc_super_a2304_ncu_full.txt",)

    def __init__(self, world=1, rank=0, global_batch=256):
        self.world, self.rank, self.global_batch = world, rank, global_batch

    @staticmethod
    def layer_input(n, c, h, lo, hi, device="cpu"):
        """Samples [lo, hi) of the synthetic input of distinct geometry `n`: every sample has its own seed, so
        any rank (and the parity check) can draw any slice of the global batch of 256."""
        gen = torch.Generator(device="cpu")
        out = torch.empty(hi - lo, c, h, h)
        for i in range(lo, hi):
            gen.manual_seed(1000 * n + i)
            out[i - lo] = torch.rand(c, h, h, generator=gen)
        return out.to(device)

    def setup(self, device, seed=0):
        from einconv_b200 import distributed as D
        from einconv_b200 import evaluator

        self.D = D
        lo, hi = D.shard_bounds(self.global_batch, self.world, self.rank)
        self.lo, self.hi = lo, hi
        self.inputs, self.specs, self.index = [], [], []
        for n, (c, h, k, s, p, count) in enumerate(RESNET50_LAYERS):
            # one tensor per distinct geometry (layers sharing it see the same synthetic input)
            torch.manual_seed(seed * 100 + n)
            x = torch.rand(hi - lo, c, h, h, device=device)
            self.inputs.append(x)
            self.specs.append(D.LayerSpec(kernel_size=k, stride=s, padding=p))
            self.index.append(count)
        # fixed inputs -> the sweep is built once: factors written in place into flat exchange buckets, one
        # CUDA graph per bucket, bucket b's collective overlapping bucket b+1's compute
        xs, specs = [], []
        for x, spec, count in zip(self.inputs, self.specs, self.index):
            xs += [x] * count
            specs += [spec] * count
        evaluator.set_fp32_math(self.fp32_math)
        self.sweep = D.FactorSweep(xs, specs, self.global_batch,
                                   use_graphs=not os.environ.get("EINCONV_BENCH_NO_GRAPHS"))

    def step(self):
        self.out = self.sweep.run()

    def work(self):
        return 106

    def flops(self, triangle=False):
        total = 0
        for c, h, k, s, p, count in RESNET50_LAYERS:
            o = (h + 2 * p - k) // s + 1
            a = c * k * k
            total += count * 2 * self.global_batch * o * o * (a * (a + 1) // 2 if triangle else a * a)
        return total

    def dominant_kernel(self, device, reps=10):
        """The KFC call of the most frequent 3x3 geometry (C=256, 14x14, k3 p1: 5 of the 16 3x3 layers, each
        532.7 GFLOP in the full-GEMM convention) timed alone with CUDA events on this rank's batch shard: the
        call is absmax + fp16 scaling + flat-plane copies + kfc_super_kernel<__half> + the finish pass; the
        committed ncu capture gives the share of the main kernel (347 of ~460 us at batch 256)."""
        from einconv_b200 import evaluator

        x = torch.rand(self.hi - self.lo, 256, 14, 14, device=device)
        for _ in range(3):
            evaluator.kfc(x, 3, padding=1)
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            evaluator.kfc(x, 3, padding=1)
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / reps
        A, pos = 2304, (self.hi - self.lo) * 196
        full = 2.0 * pos * A * A
        return dict(call="evaluator.kfc(x[B,256,14,14], 3, padding=1), fp32 tensors in TF32 mode", batch=self.hi - self.lo,
                    ms_per_call=ms, achieved_full_gemm_tflops=full / (ms * 1e-3) / 1e12,
                    achieved_triangle_tflops=full * (A + 1) / (2 * A) / (ms * 1e-3) / 1e12,
                    main_kernel="kfc_super_kernel<__half> (tcgen05 kind::f16 on power-of-two-scaled fp16 copies)",
                    timing=f"{reps} calls back to back in one CUDA-event pair")

    def e2e_setup(self, device):
        self.x_hosts = [torch.empty(x.shape, dtype=x.dtype).pin_memory().copy_(x) for x in self.inputs]
        out = self.sweep.run()
        self.f_hosts = {k: [torch.empty(f.shape, dtype=f.dtype).pin_memory() for f in fs] for k, fs in out.items()}

    def e2e_step(self, device):
        for x, xh in zip(self.inputs, self.x_hosts):
            x.copy_(xh, non_blocking=True)
        out = self.sweep.run()
        d2h = 0
        for kind, fs in out.items():
            for f, fh in zip(fs, self.f_hosts[kind]):
                fh.copy_(f, non_blocking=True)
                d2h += _nbytes(fh)
        return _nbytes(*self.x_hosts), d2h

    CHUNK = 2

    def cpu_run(self, ref):
        """the reference cannot run N=256 (24 GB of temporaries per layer, 370 GB for the
        stem); it runs in batch chunks whose factors are summed.  Timed: ONE chunk of CHUNK samples through all
        106 factors; a sweep of 256 samples is 256 / CHUNK such chunks."""
        t_total = 0.0
        self.cpu_chunk = []
        for n, (c, h, k, s, p, count) in enumerate(RESNET50_LAYERS):
            x = self.inputs[n][: self.CHUNK].cpu()
            t0 = time.perf_counter()
            kf = ref.kfc(x, k, stride=s, padding=p)
            kr = ref.kfac_reduce(x, k, stride=s, padding=p)
            t_total += (time.perf_counter() - t0) * count
            self.cpu_chunk.append((kf, kr))
        chunks = self.global_batch / self.CHUNK
        return t_total * chunks, self.work(), (
            f"one chunk of {self.CHUNK}/256 samples through all 106 factors, timed once per distinct geometry and "
            f"multiplied by the layer count, extrapolated x{chunks:.0f} chunks (the reference cannot hold N=256)")

    def chunk_parity(self):
        """The GPU factors of the chunk the CPU baseline just computed, against those CPU results
        (the chunk term of the chunk-summed oracle of SURVEY.md 8d).  Max error relative to max|ref|."""
        from einconv_b200 import evaluator

        worst = 0.0
        for n, (c, h, k, s, p, count) in enumerate(RESNET50_LAYERS):
            x = self.inputs[n][: self.CHUNK]
            for fn, want in zip((evaluator.kfc, evaluator.kfac_reduce), self.cpu_chunk[n]):
                got = fn(x, k, stride=s, padding=p).double().cpu()
                worst = max(worst, float((got - want.double()).abs().max() / want.double().abs().max()))
        return worst


WORKLOADS = dict(unfold=UnfoldC2, conv_fwd=ConvFwdC1, depthwise=DepthwiseC3, weight_vjp=WeightVjpC4,
                 factors=FactorsC5)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), reasons=sorted(reasons))


def flush_l2(buf):
    """Evict the workload's tensors from the 126 MB L2: write 256 MiB, then read 256 MiB so that the
    cache ends up holding CLEAN lines (dirty flush lines would otherwise be written back inside the
    timed region and be charged to the kernel under test)."""
    half = buf.numel() // 2
    buf[:half].zero_()
    buf[half:].sum()


def time_steps(wl, steps, warmup, device, flush=True, step=None):
    """Per-step CUDA-event timing on the current stream; returns (list of ms, launches per step)."""
    from einconv_b200 import _native as nat

    step = step or wl.step
    buf = torch.zeros(512 << 20, dtype=torch.uint8, device=device).view(torch.int32) if flush else None
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(device)
    times = []
    launches0 = nat.launch_count()
    for _ in range(steps):
        if flush:
            flush_l2(buf)
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record()
        step()
        stop.record()
        stop.synchronize()
        times.append(start.elapsed_time(stop))
    launches = (nat.launch_count() - launches0) / max(1, steps)
    if step == wl.step:
        if getattr(wl, "graph", None) is not None:
            launches += wl.graph_launches  # replayed from the workload's CUDA graph
        sweep = getattr(wl, "sweep", None)
        if sweep is not None and getattr(sweep, "_graphs", None) is not None:
            launches += sweep.launches_per_run  # kernels replayed from CUDA graphs are not counted by the library
    return times, launches


def to_value(wl, ms, n_units=1):
    """Metric value for `n_units` workloads finishing in `ms`."""
    if wl.unit == "GB/s":
        return wl.work() * n_units / (ms * 1e-3) / 1e9
    if wl.unit == "TFLOP/s":
        return wl.work() * n_units / (ms * 1e-3) / 1e12
    return wl.work() / (ms * 1e-3)  # factors/s: the sweep is shared by all ranks


def ncu_traffic(wl):
    """Sum over the workload's committed ncu captures of dram__bytes_read.sum + dram__bytes_write.sum
    (bytes per launch): every kernel of one step, not just the dominant one.  None without captures."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, sources = 0.0, []
    for rel in wl.ncu_profiles:
        if not (ROOT / rel).exists():
            return None
        seen = 0  # a summary holds one or more kernels (all of them launches of the step): read + write of each
        for line in (ROOT / rel).read_text().splitlines():
            parts = line.split()
            if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(parts[1]) * scale.get(parts[2], 1.0)
                seen += 1
        if seen < 2 or seen % 2:
            return None
        sources.append(rel)
    if not sources:
        return None
    total *= getattr(wl, "traffic_launches", 1)
    return dict(bytes=total, source=" + ".join(sources))


def roofline(wl, ms, pk, device, n_gpus=1):
    """`ms` = time of one step of ONE rank; the peak is one GPU's, so a per-rank figure goes in."""
    if wl.bound == "hbm":
        achieved = wl.work() / (ms * 1e-3) / 1e9
        r = dict(bound="hbm", achieved=achieved, peak=pk["hbm"], unit="GB/s", frac=achieved / pk["hbm"],
                 traffic=None, peak_source=pk["source"])
    else:
        # per-GPU share of the algorithmic flops (the factor sweep splits a fixed batch over the ranks)
        flops = (wl.flops() / n_gpus) if hasattr(wl, "flops") else wl.work()
        achieved = flops / (ms * 1e-3) / 1e12
        r = dict(bound="tensor", achieved=achieved, peak=pk["bf16"], unit="TFLOP/s", frac=achieved / pk["bf16"],
                 traffic=None, peak_source=pk["source"], per_gpu=True,
                 note="peak = measured dense bf16 cuBLAS of one GPU; achieved = this rank's share of the algorithmic "
                      "flops / its step time")
        if wl.fp32_math == "tf32":
            r["peak_tf32"] = tf32_peak(device)
            r["peak_tf32_source"] = "measured live in this run: torch.matmul fp32 allow_tf32 8192^3, best of 10, CUDA events"
            r["frac_tf32"] = achieved / r["peak_tf32"]
        if hasattr(wl, "flops"):
            tri = wl.flops(triangle=True) / n_gpus / (ms * 1e-3) / 1e12
            r["flop_convention"] = ("achieved counts the FULL A x A GEMM of every KFC factor (what the reference "
                                    "computes); the kernels compute only the upper triangle tiles and mirror them")
            r["achieved_triangle_flops"] = tri
            r["frac_triangle_tf32"] = tri / r["peak_tf32"]
        if hasattr(wl, "bytes_algorithmic"):
            r["algorithmic_bytes"] = wl.bytes_algorithmic()
    t = ncu_traffic(wl)
    if t:
        r["traffic"] = t["bytes"]
        r["traffic_source"] = t["source"] + " (dram__bytes_read.sum + dram__bytes_write.sum per launch, every kernel of the step)"
    return r


def config_for(wl, world):
    return dict(workload=wl.name,
                l2="flushed between steps (256 MiB write + 256 MiB read, outside the timed events)",
                timing="CUDA events per step on the launch stream, max over ranks",
                parallelism=f"batch shards x{world}, no collective" if not isinstance(wl, FactorsC5)
                else f"batch 256 split x{world}, KFC all-reduce + KFAC-reduce all-gather over NCCL")


def measure_cpu(wl, ref, repeats=3):
    """Median of `repeats` runs after one warm-up (workloads whose sample is long run once)."""
    torch.set_num_threads(os.cpu_count() or 1)
    t, units, sample = wl.cpu_run(ref)  # warm-up (builds the index patterns' caches, if any)
    if t < 5.0:
        runs = [wl.cpu_run(ref) for _ in range(repeats)]
        t = statistics.median(r[0] for r in runs)
        sample += f", median of {repeats} after 1 warm-up"
    else:
        sample += ", single run"
    if wl.unit == "GB/s":
        value = units / t / 1e9
    elif wl.unit == "TFLOP/s":
        value = units / t / 1e12
    else:
        value = units / t
    return dict(value=value, unit=wl.unit, cores=torch.get_num_threads(), kind=ref.kind,
                sample=sample + f" ({ref.source})", seconds=t)


def run_reference(args):
    """The reference's own CPU implementation on the host cores: full C2 batch, same metric and config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ReferenceCPU()
    wl = WORKLOADS[args.workload]()
    wl.setup(torch.device("cpu"))
    for _ in range(max(1, args.warmup)):
        wl.cpu_run(ref)
    times, units, sample = [], 0, ""
    for _ in range(args.steps):
        t, units, sample = wl.cpu_run(ref)
        times.append(t)
    ms = statistics.mean(times) * 1e3
    value = units / (ms * 1e-3) / (1e9 if wl.unit == "GB/s" else 1e12 if wl.unit == "TFLOP/s" else 1)
    line = dict(
        metric=wl.metric, value=value, unit=wl.unit, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        impl="reference", config=config_for(wl, args.gpus),
        cpu_baseline=dict(value=value, unit=wl.unit, cores=torch.get_num_threads(), kind=ref.kind,
                          sample=sample + f" ({ref.source})"),
        e2e=dict(value=value, unit=wl.unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="unfold", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-suite", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    import __graft_entry__

    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from einconv_b200 import _native as nat
    from einconv_b200 import evaluator

    nat.lib()
    pk = peaks()
    ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = ReferenceCPU()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    def make(name):
        cls = WORKLOADS[name]
        return cls(world, rank) if name == "factors" else cls()

    def measure(name, steps, warmup, sample_clocks=False):
        """Everything about one workload: device-timed value, roofline, e2e, eager latency, cpu baseline."""
        wl = make(name)
        evaluator.set_fp32_math(wl.fp32_math)
        wl.setup(device, seed=rank)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None
        times, launches = time_steps(wl, steps, warmup, device)
        clocks = sampler.stop() if sampler else None
        step_ms = statistics.mean(times)
        ms = max_over_ranks(step_ms)
        n_units = world if name != "factors" else 1
        res = dict(wl=wl, ms=ms, launches=launches, clocks=clocks, value=to_value(wl, ms, n_units))

        # roofline of the dominant kernel: its average launch duration (CUDA events around back-to-back
        # launches on rotating cold buffers) where the workload provides that measurement, else the step time
        kern_ms = wl.kernel_ms(device) if hasattr(wl, "kernel_ms") else None
        if kern_ms is not None:
            kern_ms = max_over_ranks(kern_ms)
        roof = roofline(wl, kern_ms if kern_ms is not None else ms, pk, device, n_gpus=world if name == "factors" else 1)
        roof["kernel_ms"] = kern_ms
        roof["timing"] = ("8 launches back to back in one CUDA-event pair, rotating input/output buffers (2.6 GB >> L2), "
                          "best of 3" if kern_ms is not None else "per-step CUDA events (includes the launch gaps)")
        if hasattr(wl, "dominant_kernel"):
            roof["dominant_kernel"] = wl.dominant_kernel(device)
            roof["dominant_kernel"]["frac_of_bf16_peak_full_gemm"] = roof["dominant_kernel"]["achieved_full_gemm_tflops"] / pk["bf16"]
            roof["dominant_kernel"]["frac_of_bf16_peak_triangle"] = roof["dominant_kernel"]["achieved_triangle_tflops"] / pk["bf16"]
            roof["traffic_note"] = "traffic = DRAM bytes of ONE launch of the dominant kernel (kfc_super_kernel on the C=256 14x14 layer, batch 256)"
        if wl.bound == "hbm":
            roof["frac_per_step"] = wl.work() / (step_ms * 1e-3) / 1e9 / pk["hbm"]
        if hasattr(wl, "naive_bytes"):
            roof["achieved_naive_bytes"] = wl.naive_bytes() / ((kern_ms or step_ms) * 1e-3) / 1e9
        res["roofline"] = roof

        # eager (no CUDA graph) latency of the same calls through the public API, L2 flushed like the steps
        if hasattr(wl, "eager_step"):
            et, _ = time_steps(wl, min(steps, 10), 3, device, step=wl.eager_step)
            res["eager_ms_per_step"] = max_over_ranks(statistics.mean(et))
            res["plan_cache"] = evaluator.plan_cache_info()
        if hasattr(wl, "also") and world == 1:
            try:
                res["also"] = wl.also(device)
            except Exception as e:
                res["also"] = dict(error=f"{type(e).__name__}: {e}")

        # ---- end to end through the public API with host buffers -------------------------------
        wl.e2e_setup(device)
        for _ in range(2):
            wl.e2e_step(device)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        n_e2e = max(3, min(steps, 10))
        t0 = time.perf_counter()
        h2d = d2h = 0
        for _ in range(n_e2e):
            h2d, d2h = wl.e2e_step(device)
            torch.cuda.synchronize(device)
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / n_e2e)
        res["e2e"] = dict(value=to_value(wl, e2e_ms, n_units), unit=wl.unit, h2d_bytes_per_step=h2d,
                          d2h_bytes_per_step=d2h, ms_per_step=e2e_ms,
                          api="pinned host tensors -> .to(device) -> public API call -> copy into pinned host tensors, "
                              "synchronised every step")

        # ---- CPU baseline (the reference's own einsum path) on rank 0, N=1 only ------------------
        res["cpu_baseline"] = None
        if ref is not None:
            try:
                res["cpu_baseline"] = measure_cpu(wl, ref)
                if hasattr(wl, "chunk_parity"):
                    err = wl.chunk_parity()
                    res["cpu_baseline"]["parity_checked"] = bool(err <= 1e-3)
                    res["cpu_baseline"]["parity_max_err_rel_to_max"] = err
                    res["cpu_baseline"]["parity_tolerance"] = 1e-3
            except Exception as e:  # a failing baseline must not lose the GPU numbers
                res["cpu_baseline"] = dict(error=f"{type(e).__name__}: {e}", kind=ref.kind)
        return res

    def factor_exchange_parity(wl):
        """N > 1: the exchanged factors of this job against a single-GPU full-batch run on rank 0 (outside any
        timed region).  Rank 0 re-draws every rank's shard (seed = rank * 100 + geometry) and evaluates the
        concatenated batch of 256 with the single-process API."""
        out = wl.sweep.run()
        torch.cuda.synchronize(device)
        worst, checked = 0.0, 0
        if rank == 0:
            layer = 0
            for n, (c, h, k, s, p, count) in enumerate(RESNET50_LAYERS):
                if n % 3 == 0 or c * k * k >= 2304:  # a third of the geometries plus every large factor
                    parts = []
                    for r in range(world):
                        lo, hi = wl.D.shard_bounds(wl.global_batch, world, r)
                        torch.manual_seed(r * 100 + n)
                        parts.append(torch.rand(hi - lo, c, h, h, device=device))
                    full = torch.cat(parts)
                    del parts
                    for kind, fn in (("kfc", evaluator.kfc), ("kfac_reduce", evaluator.kfac_reduce)):
                        want = fn(full, k, stride=s, padding=p).double()
                        got = out[kind][layer].double()
                        worst = max(worst, float((got - want).abs().max() / want.abs().max()))
                        checked += 1
                    del full
                layer += count
        return dict(parity_checked=bool(worst <= 1e-3) if rank == 0 else None, factors_compared=checked,
                    max_err_rel_to_max=worst, tolerance=1e-3,
                    against="single-GPU evaluation of the concatenated batch of 256 on rank 0")

    def factor_weak_scaling(strong_ms):
        """Companion of the strong-scaling factor number (informational; BASELINE's config is the fixed batch of
        256): every rank keeps a full batch of 256 (global batch 256 x N), same sweep, same all-reduce of every
        factor.  Separates the cost of the exchange (all that weak scaling adds to the N = 1 sweep) from the
        A x A work per factor that does not shrink with the per-rank batch under strong scaling."""
        del_wl = FactorsC5(world, rank, global_batch=256 * world)
        evaluator.set_fp32_math(del_wl.fp32_math)
        del_wl.setup(device, seed=rank)
        dist.barrier()
        torch.cuda.synchronize(device)
        times, _ = time_steps(del_wl, 5, 3, device)
        ms = max_over_ranks(statistics.mean(times))
        return dict(global_batch=256 * world, per_rank_batch=256, ms_per_step=ms, samples_per_s=256 * world / (ms * 1e-3),
                    strong_samples_per_s=256 / (strong_ms * 1e-3),
                    note="same 106 factors and the same NCCL exchange per step as the strong-scaling run")

    head = measure(args.workload, args.steps, args.warmup, sample_clocks=True)
    wl = head["wl"]
    exchange = factor_exchange_parity(wl) if (args.workload == "factors" and world > 1) else None

    # ---- the other BASELINE configs, measured the same way ---------------------------------------
    suite = {}
    if not args.no_suite:
        names = [n for n in ("conv_fwd", "depthwise", "weight_vjp", "factors") if n != args.workload]
        if args.workload != "unfold":
            names.insert(0, "unfold")
        for name in names:
            if world > 1 and name != "factors":
                continue  # only the factor path has an exchange step; the rest is replicas
            steps = min(args.steps, 10)
            try:
                torch.cuda.empty_cache()
                r = measure(name, steps, 3)
                w2 = r["wl"]
                entry = dict(workload=w2.name, metric=w2.metric, value=r["value"], unit=w2.unit, steps=steps,
                             ms_per_step=r["ms"], gpu_launches=r["launches"], dtype=w2.dtype,
                             scaling="strong" if name == "factors" else "weak",
                             roofline=r["roofline"], e2e=r["e2e"], cpu_baseline=r["cpu_baseline"])
                for key in ("eager_ms_per_step", "plan_cache", "also"):
                    if key in r:
                        entry[key] = r[key]
                if name == "factors" and world > 1:
                    entry["exchange_parity"] = factor_exchange_parity(w2)
                    entry["weak_scaling"] = factor_weak_scaling(r["ms"])
                suite[name] = entry
                del r, w2
            except Exception as e:  # keep the headline line even if a suite entry fails
                suite[name] = dict(error=f"{type(e).__name__}: {e}")
            torch.cuda.empty_cache()

    if rank == 0:
        line = dict(
            metric=wl.metric, value=head["value"], unit=wl.unit, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=head["ms"], higher_is_better=True,
            # the factor workload splits a FIXED batch of 256 over the ranks; every other workload gives
            # each rank its own full batch
            scaling="strong" if args.workload == "factors" else "weak", vs_baseline=None, dtype=wl.dtype,
            data="synthetic", config=config_for(wl, world),
            roofline=head["roofline"], cpu_baseline=head["cpu_baseline"], e2e=head["e2e"],
            gpu_launches=head["launches"], clocks=head["clocks"], suite=suite,
        )
        for key in ("eager_ms_per_step", "plan_cache", "also"):
            if key in head:
                line[key] = head[key]
        if exchange is not None:
            line["exchange_parity"] = exchange
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
