#!/usr/bin/env python
"""Benchmark of the einconv hot path on B200 (see BASELINE.md for workloads and algorithmic work).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload unfold|conv_fwd|depthwise|weight_vjp|factors]
                    [--impl ours|reference] [--no-suite]

Prints ONE JSON line.  The headline workload is BASELINE.json ``configs[1]`` (C2: unfoldNd 3-d,
N=16 C=32 16x56x56 k=3 stride=2 dilation=2, fp32, bit-exact gather) with metric "unfoldNd HBM GB/s"
= algorithmic bytes (249.97 MB per call: 224.28 MB written + 25.69 MB compulsory read) / time.
A "step" is one call over one synthetic batch.  ``value`` is timed with CUDA events on the launch
stream with inputs resident in HBM and L2 flushed between steps; ``e2e`` is the same call through
``einconv_b200.functionals`` with HOST (pinned) buffers, host<->device copies inside the timed
region.  ``suite`` carries the other BASELINE configs measured the same way (kernel-only).
With ``--gpus N`` (launched by torchrun) every rank runs the workload on its own batch shard
("weak" scaling; the factor workload adds the NCCL all-reduce of SURVEY.md 8e) and the time is the
max over ranks.  ``--impl reference`` times the reference's CPU algorithm (oracle/einsum_port.py:
one-hot patterns + torch.einsum) on the host cores for the same workload and metric.
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


PEAK_TF32_MEASURED = 734.7  # TFLOP/s


def peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        p = json.loads(path.read_text())
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# workloads: build(device, batch_scale) -> dict(run=callable, unit, work per step, ...)
# ------------------------------------------------------------------------------------------------
class Workload:
    name = ""
    metric = ""
    unit = ""
    bound = "hbm"
    dtype = "f32"
    fp32_math = "fp32"  # how fp32 contractions run: "fp32" (CUDA cores, exact) or "tf32" (tcgen05)

    def setup(self, device, seed=0):
        raise NotImplementedError

    def step(self):
        raise NotImplementedError

    def work(self):
        """Algorithmic units per step (bytes for GB/s metrics, flops for TFLOP/s, factors)."""
        raise NotImplementedError

    def host_inputs(self):
        return []

    graph = None
    graph_launches = 0

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

    def e2e_step(self, dev_inputs):
        raise NotImplementedError


class UnfoldC2(Workload):
    name = "C2 unfoldNd 3d N=16 C=32 16x56x56 k=3 stride=2 dilation=2 fp32"
    metric = "unfoldNd HBM GB/s"
    unit = "GB/s"
    bound = "hbm"
    shape = (16, 32, 16, 56, 56)
    kw = dict(kernel_size=3, dilation=2, padding=0, stride=2)

    def setup(self, device, seed=0):
        from einconv_b200.functionals import unfoldNd

        torch.manual_seed(seed)
        self.fn = unfoldNd
        self.x_host = torch.rand(*self.shape).pin_memory() if device.type == "cuda" else torch.rand(*self.shape)
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

    def e2e(self, device, out_host):
        # public host-buffer API: batch chunks alternate between two streams so that uploads and the
        # gather overlap the (dominant) download of the previous chunk
        from einconv_b200.functionals import unfoldNd_host

        unfoldNd_host(self.x_host, out=out_host, device=device, **self.kw)
        return self.x_host.numel() * 4, out_host.numel() * 4

    def e2e_out(self):
        return torch.empty(16, 864, 4056).pin_memory()

    def cpu_run(self, frac=1.0):
        from oracle import einsum_port

        n = max(1, int(round(16 * frac)))
        x = self.x_host[:n]
        t0 = time.perf_counter()
        einsum_port.unfoldNd(x, 3, dilation=2, padding=0, stride=2)
        return time.perf_counter() - t0, self.work() * n / 16, f"{n}/16 samples of the C2 batch"


class ConvFwdC1(Workload):
    name = "C1 convNd 2d fwd N=32 C=64->64 56x56 k3 s1 p1 fp32 (step = one CUDA-graph replay of the call)"
    metric = "convNd fwd TFLOP/s"
    unit = "TFLOP/s"
    bound = "tensor"

    def setup(self, device, seed=0):
        from einconv_b200.functionals import convNd

        torch.manual_seed(seed)
        self.fn = convNd
        self.x = torch.rand(32, 64, 56, 56).to(device)
        self.w = torch.rand(64, 64, 3, 3).to(device)
        self.capture(self._calls)

    dtype = "f32 (tf32 MMA, fp32 accumulate)"
    fp32_math = "tf32"

    def _calls(self):
        self.out = self.fn(self.x, self.w, padding=1)

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._calls()

    def work(self):
        return 2 * 32 * 64 * 64 * 9 * 56 * 56


class DepthwiseC3(Workload):
    name = "C3 depthwise 2d fwd + input VJP N=32 C=256 28x28 k3 p1 fp32 (step = one CUDA-graph replay of the two calls, forked streams)"
    metric = "depthwise fwd+input-VJP HBM GB/s"
    unit = "GB/s"
    bound = "hbm"

    def setup(self, device, seed=0):
        from einconv_b200 import evaluator
        from einconv_b200.functionals import convNd

        torch.manual_seed(seed)
        self.fwd, self.vjp = convNd, evaluator.conv_input_vjp_raw
        self.x = torch.rand(32, 256, 28, 28).to(device)
        self.w = torch.rand(256, 1, 3, 3).to(device)
        self.v = torch.rand(32, 256, 28, 28).to(device)
        self.side = None
        self.capture(self._calls)

    def _calls(self):
        # the two directions are independent: the input VJP runs on a forked stream, so that inside the
        # captured graph the two ~12 us kernels overlap instead of queueing behind each other
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

    def work(self):
        one = 2 * 32 * 256 * 28 * 28 * 4 + 256 * 9 * 4
        return 2 * one


class WeightVjpC4(Workload):
    name = "C4 convNd 3d weight VJP N=8 C=64 16x112x112 k3 p1 bf16"
    metric = "convNd weight-VJP TFLOP/s"
    unit = "TFLOP/s"
    bound = "tensor"
    dtype = "bf16"

    def setup(self, device, seed=0):
        from einconv_b200 import evaluator

        torch.manual_seed(seed)
        self.fn = evaluator.conv_weight_vjp_raw
        self.x = torch.rand(8, 64, 16, 112, 112).bfloat16().to(device)
        self.v = torch.rand(8, 64, 16, 112, 112).bfloat16().to(device)

    def step(self):
        self.out = self.fn(self.x, self.v, 3, 1, 1, 1, 1)

    def work(self):
        return 2 * 8 * 64 * 64 * 27 * 16 * 112 * 112


class FactorsC5(Workload):
    """KFC + KFAC-reduce over the 53 ResNet-50 conv layers, batch 256 split over the ranks."""

    name = "C5 KFC + KFAC-reduce, 53 ResNet-50 conv layers, batch 256 (sharded by N), fp32"
    metric = "KFC+KFAC-reduce factors/s"
    unit = "factors/s"
    bound = "tensor"
    dtype = "f32 (KFC: tf32 MMA, fp32 accumulate; KFAC-reduce: fp32)"
    fp32_math = "tf32"

    def __init__(self, world=1, rank=0, global_batch=256):
        self.world, self.rank, self.global_batch = world, rank, global_batch

    def setup(self, device, seed=0):
        from einconv_b200 import distributed as D

        self.D = D
        lo, hi = D.shard_bounds(self.global_batch, self.world, self.rank)
        self.local = hi - lo
        self.inputs, self.specs, self.index = [], [], []
        for n, (c, h, k, s, p, count) in enumerate(RESNET50_LAYERS):
            torch.manual_seed(seed + n)
            # one tensor per distinct geometry (layers sharing it see the same synthetic input)
            x = torch.rand(self.local, c, h, h, device=device)
            self.inputs.append(x)
            self.specs.append(D.LayerSpec(kernel_size=k, stride=s, padding=p))
            self.index.append(count)

        # fixed inputs -> the sweep is built once: factors written in place into flat all-reduce
        # buckets, one CUDA graph per bucket, bucket b's all-reduce overlapping bucket b+1's compute
        xs, specs = [], []
        for x, spec, count in zip(self.inputs, self.specs, self.index):
            xs += [x] * count
            specs += [spec] * count
        from einconv_b200 import evaluator

        evaluator.set_fp32_math(self.fp32_math)
        self.sweep = D.FactorSweep(xs, specs, self.global_batch, use_graphs=not os.environ.get("EINCONV_BENCH_NO_GRAPHS"))

    def step(self):
        self.out = self.sweep.run()

    def work(self):
        return 106

    def flops(self):
        total = 0
        for c, h, k, s, p, count in RESNET50_LAYERS:
            o = (h + 2 * p - k) // s + 1
            a = c * k * k
            total += count * 2 * self.global_batch * o * o * a * a
        return total


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


def time_steps(wl, steps, warmup, device, flush=True):
    """Per-step CUDA-event timing on the current stream; returns (list of ms, launches per step)."""
    from einconv_b200 import _native as nat

    buf = torch.zeros(512 << 20, dtype=torch.uint8, device=device).view(torch.int32) if flush else None
    for _ in range(warmup):
        wl.step()
    torch.cuda.synchronize(device)
    times = []
    launches0 = nat.launch_count()
    for _ in range(steps):
        if flush:
            flush_l2(buf)
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record()
        wl.step()
        stop.record()
        stop.synchronize()
        times.append(start.elapsed_time(stop))
    launches = (nat.launch_count() - launches0) / max(1, steps)
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


# ncu --set full captures of the dominant kernel of each workload (summaries committed under profiles/)
NCU_PROFILES = {
    "unfold": "profiles/r01_unfold_c2_ncu_full_v10_tma.txt",
    "depthwise": "profiles/r01_depthwise_c3_ncu_full_v5_vec4_nr6.txt",
    "conv_fwd": "profiles/r01_conv_fwd_c1_ncu_full_v3.txt",
    "weight_vjp": "profiles/r01_wgrad_c4_ncu_full_v3_quad.txt",
}


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the workload's dominant kernel,
    read from the committed ncu summary; None if there is no capture."""
    rel = NCU_PROFILES.get(workload)
    if not rel or not (ROOT / rel).exists():
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, 0
    for line in (ROOT / rel).read_text().splitlines():
        parts = line.split()
        if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and seen < 2:
            total += float(parts[1]) * scale.get(parts[2], 1.0)
            seen += 1
    return dict(bytes=total, source=rel) if seen == 2 else None


def roofline(wl, ms_kernel, pk):
    r = _roofline(wl, ms_kernel, pk)
    key = next((k for k, c in WORKLOADS.items() if isinstance(wl, c)), None)
    t = ncu_traffic(key)
    if t:
        r["traffic"] = t["bytes"]
        r["traffic_source"] = t["source"] + " (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
    return r


def _roofline(wl, ms_kernel, pk):
    if wl.bound == "hbm":
        achieved = wl.work() / (ms_kernel * 1e-3) / 1e9
        return dict(bound="hbm", achieved=achieved, peak=pk["hbm"], unit="GB/s", frac=achieved / pk["hbm"],
                    traffic=None, peak_source=pk["source"])
    flops = wl.flops() if hasattr(wl, "flops") else wl.work()
    achieved = flops / (ms_kernel * 1e-3) / 1e12
    r = dict(bound="tensor", achieved=achieved, peak=pk["bf16"], unit="TFLOP/s", frac=achieved / pk["bf16"],
             traffic=None, peak_source=pk["source"],
             note="peak = measured dense bf16 cuBLAS; fp32 runs on CUDA cores / tf32 at half that rate")
    if wl.fp32_math == "tf32":
        # dense TF32 cuBLAS on this pool's B200 (scripts/measure_tf32_peak.py, 4096^3, best of 5 x 10)
        r["peak_tf32"] = PEAK_TF32_MEASURED
        r["frac_tf32"] = achieved / PEAK_TF32_MEASURED
    return r


def run_reference(args):
    """The reference's CPU algorithm (oracle port) on the host cores, same workload and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    wl = UnfoldC2()
    wl.setup(torch.device("cpu"))
    frac = 0.25  # bounded sample: 4 of the 16 batch elements per step
    for _ in range(max(1, args.warmup)):
        wl.cpu_run(frac)
    times, work, sample = [], 0, ""
    for _ in range(args.steps):
        t, work, sample = wl.cpu_run(frac)
        times.append(t)
    ms = statistics.mean(times) * 1e3
    value = work / (ms * 1e-3) / 1e9
    line = dict(
        metric=wl.metric, value=value, unit=wl.unit, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        impl="reference", config=dict(workload=wl.name, sample=sample),
        cpu_baseline=dict(value=value, unit=wl.unit, cores=torch.get_num_threads(), kind="port", sample=sample),
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

    nat.lib()
    pk = peaks()

    def make(name):
        cls = WORKLOADS[name]
        return cls(world, rank) if name == "factors" else cls()

    def measure(name, steps, warmup):
        from einconv_b200 import evaluator

        wl = make(name)
        evaluator.set_fp32_math(wl.fp32_math)
        wl.setup(device, seed=rank)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        sampler = ClockSampler(local_rank) if rank == 0 else None
        times, launches = time_steps(wl, steps, warmup, device)
        clocks = sampler.stop() if sampler else None
        ms = statistics.mean(times)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return wl, ms, launches, clocks, times

    wl, ms, launches, clocks, times = measure(args.workload, args.steps, args.warmup)
    n_units = world if args.workload != "factors" else 1
    value = to_value(wl, ms, n_units)
    # roofline of the dominant kernel: its average launch duration (CUDA events around back-to-back
    # launches on rotating cold buffers) where the workload provides that measurement, else the per-step time
    step_ms = statistics.mean(times)
    kern_ms = wl.kernel_ms(device) if hasattr(wl, "kernel_ms") else None
    if kern_ms is not None and world > 1:
        t = torch.tensor([kern_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kern_ms = float(t)
    roof = roofline(wl, kern_ms if kern_ms is not None else step_ms, pk)
    roof["kernel_ms"] = kern_ms
    roof["timing"] = ("8 launches back to back in one CUDA-event pair, rotating input/output buffers (2.6 GB >> L2), "
                      "best of 3" if kern_ms is not None else "per-step CUDA events (includes the launch gap)")
    roof["frac_per_step"] = wl.work() / (step_ms * 1e-3) / 1e9 / pk["hbm"] if wl.bound == "hbm" else None
    if hasattr(wl, "naive_bytes"):
        roof["achieved_naive_bytes"] = wl.naive_bytes() / ((kern_ms or step_ms) * 1e-3) / 1e9

    # ---- end to end through the public API with host buffers --------------------------------
    e2e = None
    if hasattr(wl, "e2e"):
        out_host = wl.e2e_out()
        for _ in range(2):
            wl.e2e(device, out_host)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        n_e2e = max(3, min(args.steps, 10))
        for _ in range(n_e2e):
            h2d, d2h = wl.e2e(device, out_host)
            torch.cuda.synchronize(device)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        if world > 1:
            t = torch.tensor([e2e_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t)
        e2e = dict(value=to_value(wl, e2e_ms, n_units), unit=wl.unit, h2d_bytes_per_step=h2d,
                   d2h_bytes_per_step=d2h, ms_per_step=e2e_ms)

    # ---- CPU baseline (oracle port = the reference's algorithm) on rank 0, N=1 only ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and hasattr(wl, "cpu_run"):
        torch.set_num_threads(os.cpu_count() or 1)
        wl.cpu_run(0.25)
        runs = [wl.cpu_run(0.25) for _ in range(3)]
        t = statistics.median(r[0] for r in runs)
        cpu = dict(value=runs[0][1] / t / 1e9, unit=wl.unit, cores=torch.get_num_threads(), kind="port",
                   sample=runs[0][2] + ", median of 3 after 1 warm-up (oracle/einsum_port.py)")

    # ---- the other BASELINE configs, kernel-only, same timing rules ----------------------------
    suite = {}
    if not args.no_suite:
        names = [n for n in ("conv_fwd", "depthwise", "weight_vjp", "factors") if n != args.workload]
        if args.workload != "unfold":
            names.insert(0, "unfold")
        for name in names:
            if world > 1 and name != "factors":
                continue  # only the factor path has an exchange step; the rest is replicas
            steps = 3 if name == "factors" else min(args.steps, 10)
            try:
                w2, ms2, l2, _, t2 = measure(name, steps, 3)
                entry = dict(workload=w2.name, metric=w2.metric, value=to_value(w2, ms2), unit=w2.unit,
                             ms_per_step=ms2, gpu_launches=l2, dtype=w2.dtype,
                             roofline=roofline(w2, statistics.mean(t2), pk))
                suite[name] = entry
            except Exception as e:  # keep the headline line even if a suite entry fails
                suite[name] = dict(error=f"{type(e).__name__}: {e}")
            torch.cuda.empty_cache()

    if rank == 0:
        line = dict(
            metric=wl.metric, value=value, unit=wl.unit, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms, higher_is_better=True,
            # the factor workload splits a FIXED batch of 256 over the ranks; every other workload gives
            # each rank its own full batch
            scaling="strong" if args.workload == "factors" else "weak", vs_baseline=None, dtype=wl.dtype,
            data="synthetic",
            config=dict(workload=wl.name, l2="flushed between steps (256 MiB write + 256 MiB read, outside the timed events)",
                        timing="CUDA events per step on the launch stream, max over ranks",
                        parallelism=f"batch shards x{world}, no collective" if args.workload != "factors"
                        else f"batch 256 split x{world}, NCCL all-reduce of factors"),
            roofline=roof, cpu_baseline=cpu, e2e=e2e, gpu_launches=launches, clocks=clocks, suite=suite,
        )
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
