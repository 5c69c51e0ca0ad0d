"""GPU parity AT the benchmarked configurations, IN the benchmarked math mode.

``bench.py`` times C1 and C5 with ``set_fp32_math("tf32")`` and C4 in bf16 at full size; the case
lists of ``test_gpu_parity.py`` are the reference's (small) ones.  Here the same calls run at
BASELINE.json's sizes against an fp64 truth computed on the GPU from the same input bits, plus
the plain-C oracle on a slice the oracle finishes in seconds.

Tolerances (stated per test): TF32 contractions ``1e-3 * max|ref|`` (10 explicit mantissa bits on the
operands, fp32 accumulation); bf16 output ``2^-7 * max|ref|`` (one bf16 rounding of an fp32 sum).
"""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from einconv_b200 import _native as nat
from einconv_b200 import evaluator
from einconv_b200.functionals import convNd, unfoldNd
from oracle import direct

pytestmark = pytest.mark.gpu
DEV = "cuda"

# (C_in, H_in, k, stride, padding) of the 19 distinct ResNet-50 conv geometries
RESNET50_GEOMETRIES = [
    (3, 224, 7, 2, 3), (64, 56, 1, 1, 0), (64, 56, 3, 1, 1), (256, 56, 1, 1, 0), (128, 56, 3, 2, 1),
    (256, 56, 1, 2, 0), (128, 28, 1, 1, 0), (128, 28, 3, 1, 1), (512, 28, 1, 1, 0), (256, 28, 3, 2, 1),
    (512, 28, 1, 2, 0), (256, 14, 1, 1, 0), (256, 14, 3, 1, 1), (1024, 14, 1, 1, 0), (512, 14, 3, 2, 1),
    (1024, 14, 1, 2, 0), (512, 7, 1, 1, 0), (512, 7, 3, 1, 1), (2048, 7, 1, 1, 0),
]


@pytest.fixture
def tf32_mode():
    previous = evaluator.get_fp32_math()
    evaluator.set_fp32_math("tf32")
    yield
    evaluator.set_fp32_math(previous)


def rel_to_max(got, want, tol, what):
    got = got.detach().double()
    want = want.to(got.device).double()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    assert err <= tol * ref, f"{what}: max |d| = {err:.3e} > {tol:.1e} * max|ref| = {tol * ref:.3e}"
    return err / max(ref, 1e-300)


# ---------------------------------------------------------------------------------------------
# C1: conv2d forward N=32 C=64->64 56x56 k3 s1 p1, fp32 tensors in TF32 mode (what bench.py times)
# ---------------------------------------------------------------------------------------------
def test_c1_full_batch_tf32_mode(tf32_mode):
    torch.manual_seed(0)
    x = torch.rand(32, 64, 56, 56, device=DEV)
    w = torch.rand(64, 64, 3, 3, device=DEV)
    b = torch.rand(64, device=DEV)
    g = evaluator._geom(32, 1, 64, 64, (56, 56), (3, 3), 1, 1, 1)[0]
    assert evaluator.selected_kernel(nat.OP_CONV_FWD, g, torch.float32) == "tcgen05_rows_conv_fwd"
    before = nat.launch_count()
    y = convNd(x, w, b, padding=1)
    assert nat.launch_count() > before
    truth = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    rel_to_max(y, truth, 1e-3, "C1 forward, TF32 mode")
    # input VJP of the same layer, same mode
    v = torch.rand_like(y)
    gx = evaluator.conv_input_vjp_raw(w, v, (56, 56), 1, 1, 1, 1)
    truth_gx = F.conv_transpose2d(v.double(), w.double(), padding=1)
    rel_to_max(gx, truth_gx, 1e-3, "C1 input VJP, TF32 mode")
    # the oracle itself on one sample (plain C, fp64)
    want = direct.conv_forward(x[:1].cpu().numpy(), w.cpu().numpy(), b.cpu().numpy(), stride=1, padding=1,
                               dilation=1, groups=1)
    rel_to_max(y[:1], torch.from_numpy(want), 1e-3, "C1 forward vs oracle, sample 0")


# ---------------------------------------------------------------------------------------------
# C4: conv3d weight VJP N=8 C=64 16x112x112 k3 p1 bf16, FULL size (split-K over 1.6 M positions)
# ---------------------------------------------------------------------------------------------
def _weight_vjp_truth_fp64(x, v, k, pad):
    """gW[co, ci, kd, kh, kw] = sum_{b, o} x[b, ci, o + k - p] v[b, co, o] in fp64, one tap at a time from
    shifted slices (stride 1, dilation 1): the definition of expressions/convNd_weight_vjp.py:51-58."""
    B, C = x.shape[:2]
    Co = v.shape[1]
    sizes = x.shape[2:]
    gw = torch.zeros(Co, C, k, k, k, dtype=torch.float64, device=x.device)
    for b in range(B):
        xb = F.pad(x[b].double(), (pad,) * 6)
        vb = v[b].double().reshape(Co, -1)
        for kd in range(k):
            for kh in range(k):
                for kw in range(k):
                    xs = xb[:, kd:kd + sizes[0], kh:kh + sizes[1], kw:kw + sizes[2]].reshape(C, -1)
                    gw[:, :, kd, kh, kw] += vb @ xs.T
    return gw


def test_c4_full_size_bf16():
    torch.manual_seed(0)
    x = torch.rand(8, 64, 16, 112, 112).bfloat16().to(DEV)
    v = torch.rand(8, 64, 16, 112, 112).bfloat16().to(DEV)
    g = evaluator._geom(8, 1, 64, 64, (16, 112, 112), (3, 3, 3), 1, 1, 1)[0]
    assert evaluator.selected_kernel(nat.OP_CONV_WEIGHT_VJP, g, torch.bfloat16) == "tcgen05_slab_wgrad"
    before = nat.launch_count()
    gw = evaluator.conv_weight_vjp_raw(x, v, 3, 1, 1, 1, 1)
    assert nat.launch_count() > before and gw.dtype == torch.bfloat16 and gw.shape == (64, 64, 3, 3, 3)
    truth = _weight_vjp_truth_fp64(x, v, 3, 1)
    # every entry is a sum of ~1.6 M positive products (~4e5): errors are one bf16 output rounding
    rel = rel_to_max(gw, truth, 2.0**-7, "C4 weight VJP, full size, bf16")
    assert rel > 0.0  # a bf16 result cannot equal the fp64 truth bit for bit: guards a vacuous comparison


@pytest.mark.parametrize("route", ["selfpack", "packed", "fused"])
def test_c4_mid_size_every_slab_route(route, monkeypatch):
    """A C4-like tensor small enough for a quick fp64 truth but long enough for several packing chunks:
    the self-packing kernel (packed copies written by its own producer warps, chunk flags), the pre-packed
    TMA route with contiguous K splits (the default deals k-blocks round-robin; test_c4_full_size_bf16 runs
    that) and the fused shared-memory producers must agree with fp64."""
    if route == "selfpack":
        monkeypatch.setenv("EINCONV_WGRAD_SELFPACK", "1")
    elif route == "packed":
        monkeypatch.setenv("EINCONV_WGRAD_CONTIGUOUS_SPLITS", "1")   # the default route with the older k-block order
    else:
        monkeypatch.setenv("EINCONV_WGRAD_FUSED", "1")
    torch.manual_seed(3)
    x = (torch.rand(4, 64, 8, 40, 40) - 0.4).bfloat16().to(DEV)
    v = (torch.rand(4, 64, 8, 40, 40) - 0.5).bfloat16().to(DEV)
    gw = evaluator.conv_weight_vjp_raw(x, v, 3, 1, 1, 1, 1)
    truth = _weight_vjp_truth_fp64(x, v, 3, 1)
    rel_to_max(gw, truth, 2.0**-7, f"C4-like mid size, {route}")


def test_c4_slice_against_the_c_oracle_bf16():
    """One sample, reduced depth/height of the C4 tensor against oracle/einconv_oracle.c (fp64, plain loops)."""
    torch.manual_seed(0)
    x = torch.rand(1, 64, 3, 28, 112).bfloat16()
    v = torch.rand(1, 64, 3, 28, 112).bfloat16()
    want = direct.conv_weight_vjp(x.double().numpy(), v.double().numpy(), 3, dilation=1, padding=1, stride=1, groups=1)
    gw = evaluator.conv_weight_vjp_raw(x.to(DEV), v.to(DEV), 3, 1, 1, 1, 1)
    rel_to_max(gw, torch.from_numpy(want), 2.0**-7, "C4 slice vs oracle")


# ---------------------------------------------------------------------------------------------
# C5: every distinct ResNet-50 geometry at the 8-rank shard size (batch 32), TF32 mode
# ---------------------------------------------------------------------------------------------
def _factor_truth_fp64(x, k, s, p, kind, global_batch, chunk=8):
    """Reference definition (test_convNd_kfc.py:47-52, test_convNd_kfac_reduce.py:48-53): unfold (bit-exact,
    proven in test_gpu_parity.py), then the outer products in fp64, accumulated over batch chunks."""
    A = x.shape[1] * k * k
    acc = torch.zeros(A, A, dtype=torch.float64, device=x.device)
    for lo in range(0, x.shape[0], chunk):
        u = unfoldNd(x[lo:lo + chunk], k, stride=s, padding=p).double()
        if kind == "kfc":
            acc += torch.einsum("nik,njk->ij", u, u)
        else:
            m = u.mean(-1)
            acc += m.T @ m
    return (acc / global_batch).unsqueeze(0)


@pytest.mark.parametrize("geometry", RESNET50_GEOMETRIES, ids=lambda g: "C{}H{}k{}s{}p{}".format(*g))
def test_c5_resnet_geometries_batch32_tf32_mode(geometry, tf32_mode):
    c, h, k, s, p = geometry
    torch.manual_seed(c + h + k)
    x = torch.rand(32, c, h, h, device=DEV)
    o = (h + 2 * p - k) // s + 1
    # a rank of the 8-GPU job scales with the GLOBAL batch 256 (einconv_b200/distributed.py)
    scale_kfc = float(torch.tensor([1.0 / 256]))
    scale_red = float(torch.tensor([1.0 / (256 * (o * o) ** 2)]))
    for kind, fn, scale in (("kfc", evaluator.kfc, scale_kfc), ("kfac_reduce", evaluator.kfac_reduce, scale_red)):
        before = nat.launch_count()
        got = fn(x, k, stride=s, padding=p, scale=scale)
        assert nat.launch_count() > before
        truth = _factor_truth_fp64(x, k, s, p, kind, 256)
        rel_to_max(got, truth, 1e-3, f"{kind} {geometry} batch 32, TF32 mode")
        assert torch.equal(got, got.transpose(1, 2)) or torch.allclose(got, got.transpose(1, 2), rtol=1e-6, atol=0)


@pytest.mark.parametrize("geometry", [(64, 56, 3, 1, 1), (512, 7, 3, 1, 1), (1024, 14, 1, 1, 0), (3, 224, 7, 2, 3)],
                         ids=lambda g: "C{}H{}k{}s{}p{}".format(*g))
def test_c5_factor_sweep_matches_direct_calls_tf32_mode(geometry, tf32_mode):
    """FactorSweep (buckets, CUDA graphs — what bench.py replays) against the fp64 truth, batch 32."""
    from einconv_b200 import distributed as D

    c, h, k, s, p = geometry
    torch.manual_seed(7)
    x = torch.rand(32, c, h, h, device=DEV)
    sweep = D.FactorSweep([x], [D.LayerSpec(k, stride=s, padding=p)], global_batch=256)
    out = sweep.run()
    torch.cuda.synchronize()
    for kind in ("kfc", "kfac_reduce"):
        truth = _factor_truth_fp64(x, k, s, p, kind, 256)
        rel_to_max(out[kind][0], truth, 1e-3, f"sweep {kind} {geometry}")


def test_c5_chunk_sum_matches_cpu_oracle_port():
    """the GPU factor of a batch equals the chunk-summed CPU einsum-path result
    (oracle/einsum_port.py on 2-sample chunks, fp64 accumulation of the chunk sums)."""
    from oracle import einsum_port

    torch.manual_seed(3)
    x = torch.rand(4, 64, 14, 14)
    acc_k = torch.zeros(1, 576, 576, dtype=torch.float64)
    acc_r = torch.zeros(1, 576, 576, dtype=torch.float64)
    for lo in range(0, 4, 2):
        acc_k += einsum_port.evaluate(einsum_port.kfc_expression(x[lo:lo + 2], 3, padding=1)).double() * (2 / 4)
        acc_r += einsum_port.evaluate(einsum_port.kfac_reduce_expression(x[lo:lo + 2], 3, padding=1)).double() * (2 / 4)
    got_k = evaluator.kfc(x.to(DEV), 3, padding=1)
    got_r = evaluator.kfac_reduce(x.to(DEV), 3, padding=1)
    np.testing.assert_allclose(got_k.cpu().double().numpy(), acc_k.numpy(), rtol=5e-5, atol=1e-6)
    np.testing.assert_allclose(got_r.cpu().double().numpy(), acc_r.numpy(), rtol=5e-5, atol=1e-9)
