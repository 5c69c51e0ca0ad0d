"""GPU parity: the sm_100a kernels (through the Python host API -> C ABI) against the oracle.

Bars (stated per test):
* unfoldNd / fold-free gathers: bit-exact (``torch.equal``) for fp32, fp64, bf16, fp16;
* fp32 contractions in EINCONV_MATH_FP32 (default): the reference's own tolerances
  (``test/utils.py:10-16`` rtol 1e-5 / small atol; factors rtol 5e-5, ``test_convNd_kfc.py:59``);
* fp64: rtol 1e-10;
* bf16 / fp16 I/O (fp32 accumulate): compared with the fp64 oracle on the SAME rounded inputs,
  ``|delta| <= 2^-8 |ref| + small`` (bf16) resp. ``2^-11`` (fp16): one output rounding.
Every case runs with ``simplify`` True and False (specialised vs generic kernel), like the
reference's ``SIMPLIFIES`` axis (``test/utils.py:98``).
"""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import einconv_b200
from einconv_b200 import _native as nat
from einconv_b200 import evaluator
from einconv_b200.expressions import (
    convNd_forward,
    convNd_input_vjp,
    convNd_kfac_reduce,
    convNd_kfc,
    convNd_unfold,
    convNd_weight_vjp,
)
from einconv_b200.functionals import convNd, unfoldNd
from einconv_b200.modules import ConvNd, UnfoldNd
from oracle import direct
from tests import cases as C
from tests.conftest import forward_out_shape, golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
SIMPLIFY = [True, False]


def close(got, want, rtol, atol, what=""):
    got = got.detach().double().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.abs(got - want)
    bound = atol + rtol * np.abs(want)
    if not (err <= bound).all():
        idx = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(
            f"{what}: max violation at {idx}: got {got[idx]!r} want {want[idx]!r} "
            f"(|d|={err[idx]:.3e}, allowed {bound[idx]:.3e}); mismatches {(err > bound).sum()}/{err.size}"
        )


def test_native_library_is_loaded_and_counts_launches():
    before = nat.launch_count()
    x = torch.rand(1, 2, 8, 8, device=DEV)
    unfoldNd(x, 3)
    assert nat.launch_count() > before
    assert nat.LIB_PATH.exists()


# =============================================================================================
# unfold: bit-exact
# =============================================================================================
@pytest.fixture(params=["auto", "plane", "tiled"])
def unfold_variant(request, monkeypatch):
    """All unfold kernels: "auto" prefers the persistent double-buffered whole-plane kernel (4/8-byte
    types, enough work), "plane" the one-CTA-per-plane(s) kernel, "tiled" the row-tiled kernel
    (csrc/unfold.cu reads the variables per call)."""
    monkeypatch.setenv("EINCONV_UNFOLD_TILED", "1" if request.param == "tiled" else "0")
    monkeypatch.delenv("EINCONV_UNFOLD_NO_PERSISTENT", raising=False)
    monkeypatch.delenv("EINCONV_UNFOLD_PERSISTENT", raising=False)
    if request.param == "plane":
        monkeypatch.setenv("EINCONV_UNFOLD_NO_PERSISTENT", "1")
    elif request.param == "auto":
        monkeypatch.setenv("EINCONV_UNFOLD_PERSISTENT", "1")  # also where the heuristic would not pick it
    return request.param


@pytest.mark.parametrize("case", C.UNFOLD_CASES, ids=[c["id"] for c in C.UNFOLD_CASES])
@pytest.mark.parametrize("simplify", SIMPLIFY)
def test_unfold_bit_exact(case, simplify, unfold_variant):
    x = C.draw_input(case)
    want = direct.unfold(x.numpy(), case["kernel_size"], **case["kwargs"])
    got = unfoldNd(x.to(DEV), case["kernel_size"], **case["kwargs"], simplify=simplify)
    assert got.dtype == x.dtype and tuple(got.shape) == want.shape
    assert torch.equal(got.cpu(), torch.from_numpy(want))
    golden("unfold").check(case["id"], got.cpu().numpy(), 0, 0, exact=True)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float64])
@pytest.mark.parametrize("case", [C.UNFOLD_CASES[5], C.UNFOLD_CASES[11], C.UNFOLD_CASES[18]],
                         ids=lambda c: c["id"])
def test_unfold_bit_exact_other_dtypes(case, dtype, unfold_variant):
    x = C.draw_input(case).to(dtype)
    raw = x.view(torch.int16).numpy() if dtype in (torch.bfloat16, torch.float16) else x.numpy()
    want = direct.unfold(raw, case["kernel_size"], **case["kwargs"])
    got = unfoldNd(x.to(DEV), case["kernel_size"], **case["kwargs"]).cpu()
    got_raw = got.view(torch.int16).numpy() if dtype in (torch.bfloat16, torch.float16) else got.numpy()
    assert np.array_equal(got_raw, want)


UNFOLD_EXTRA = [
    dict(x=(1, 1, 11), kernel_size=4, kwargs=dict(padding=6)),                      # padding > kernel
    dict(x=(2, 2, 91), kernel_size=4, kwargs=dict(stride=2, padding=20, dilation=3)),
    dict(x=(2, 3, 9, 8), kernel_size=4, kwargs=dict(padding="same")),               # asymmetric same
    dict(x=(3, 2, 7, 7), kernel_size=7, kwargs=dict(padding=3)),
    dict(x=(2, 3, 224, 224), kernel_size=7, kwargs=dict(stride=2, padding=3)),      # row-tiled staging
    dict(x=(5, 70, 7, 7), kernel_size=3, kwargs=dict(padding=1)),                   # many tiny planes
    dict(x=(2, 2, 6, 5, 4, 7), kernel_size=(2, 3, 2, 3), kwargs=dict(padding=1, stride=(1, 2, 1, 2))),  # N = 4
    dict(x=(1, 2, 3, 4, 5, 4, 6), kernel_size=2, kwargs=dict(dilation=(1, 2, 1, 1, 2))),              # N = 5
    dict(x=(0, 3, 8, 8), kernel_size=3, kwargs={}),                                  # empty batch
    dict(x=(2, 3, 4, 4), kernel_size=5, kwargs=dict(padding=0), expect_error=True),  # kernel larger than input
]


# planes of > 64 KB of output and > 1024 positions: the whole-plane kernel stages them with ONE bulk
# tensor copy (element strides = gcd(stride, dilation), zero fill = padding) when the tensor is 16-byte
# aligned and contiguous; "software" forces the staging loop on the same shapes
UNFOLD_TMA = [
    dict(x=(2, 3, 48, 48), kernel_size=3, kwargs=dict(padding=1)),                    # start coordinate -4, 3 slack columns
    dict(x=(1, 2, 100, 104), kernel_size=3, kwargs=dict(stride=2, dilation=2)),      # rows compacted by the copy
    dict(x=(2, 2, 10, 24, 28), kernel_size=3, kwargs=dict(stride=2, dilation=2)),    # N = 3 like C2
    dict(x=(1, 2, 224, 224), kernel_size=7, kwargs=dict(stride=2, padding=3)),       # cut along the outermost dim
    dict(x=(2, 3, 4000), kernel_size=5, kwargs=dict(padding=2)),                     # window wider than a TMA box
    dict(x=(2, 3, 46, 46), kernel_size=3, kwargs=dict(padding=1)),                   # rows not 16-byte multiples
    dict(x=(1, 2, 90, 200), kernel_size=(3, 2), kwargs=dict(stride=(1, 3), dilation=(2, 6), padding=(1, 2))),  # gcd 3, two compaction batches
    dict(x=(1, 2, 60, 64), kernel_size=(5, 2), kwargs=dict(padding=(2, 5), dilation=(1, 3))),  # padding > 4
    dict(x=(2, 1, 3, 3, 40, 44), kernel_size=(2, 2, 3, 3), kwargs=dict(padding=(0, 1, 1, 1))),  # N = 4: rank-5 map
]


@pytest.mark.parametrize("staging", ["tma", "tma-compacted", "software"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64], ids=["f32", "bf16", "f64"])
@pytest.mark.parametrize("case", UNFOLD_TMA, ids=lambda c: "x".join(map(str, c["x"])) + f"-k{c['kernel_size']}")
def test_unfold_bulk_copy_staging(case, dtype, staging, monkeypatch):
    """Bit-exact on arbitrary bit patterns (NaN payloads, infinities, denormals are copied)."""
    if staging == "software":
        monkeypatch.setenv("EINCONV_UNFOLD_NO_TMA", "1")
    elif staging == "tma-compacted":
        monkeypatch.setenv("EINCONV_UNFOLD_COMPACT", "1")  # in-place compaction of strided windows after the copy
    monkeypatch.setenv("EINCONV_UNFOLD_NO_PERSISTENT", "1")
    torch.manual_seed(3)
    itype = {torch.float32: torch.int32, torch.bfloat16: torch.int16, torch.float64: torch.int64}[dtype]
    info = torch.iinfo(itype)
    bits = torch.randint(info.min, info.max, case["x"], dtype=itype)
    bits[bits == 0] = 1  # the oracle pads with integer 0 = +0.0
    want = direct.unfold(bits.numpy(), case["kernel_size"], **case["kwargs"])
    x = bits.view(dtype).to(DEV)
    got = unfoldNd(x, case["kernel_size"], **case["kwargs"])
    assert tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().view(itype).numpy(), want)
    # a view that starts one element into an allocation is not 16-byte aligned: software staging, same result
    flat = torch.zeros(x.numel() + 1, dtype=dtype, device=DEV)
    shifted = flat[1:].view(*case["x"])
    shifted.copy_(x)
    got2 = unfoldNd(shifted, case["kernel_size"], **case["kwargs"])
    assert np.array_equal(got2.cpu().view(itype).numpy(), want)


@pytest.mark.parametrize("case", UNFOLD_EXTRA, ids=lambda c: "x".join(map(str, c["x"])) + f"-k{c['kernel_size']}")
def test_unfold_edge_cases(case, unfold_variant):
    torch.manual_seed(1)
    x = torch.rand(*case["x"])
    if case.get("expect_error"):
        with pytest.raises(RuntimeError, match="can't be greater"):
            unfoldNd(x.to(DEV), case["kernel_size"], **case["kwargs"])
        return
    want = direct.unfold(x.numpy(), case["kernel_size"], **case["kwargs"])
    got = unfoldNd(x.to(DEV), case["kernel_size"], **case["kwargs"])
    assert tuple(got.shape) == want.shape
    assert torch.equal(got.cpu(), torch.from_numpy(want))


def test_unfold_non_contiguous_input_and_module(unfold_variant):
    torch.manual_seed(0)
    base = torch.rand(2, 12, 6, 10, device=DEV)
    x = base.permute(0, 2, 1, 3)[:, :, ::2, 1:9]  # [2, 6, 6, 8], strided view
    assert not x.is_contiguous()
    got = UnfoldNd((3, 2), dilation=1, padding=1, stride=2)(x)
    want = direct.unfold(x.contiguous().cpu().numpy(), (3, 2), padding=1, stride=2)
    assert torch.equal(got.cpu(), torch.from_numpy(want))
    assert torch.equal(got, F.unfold(x.contiguous(), (3, 2), padding=1, stride=2))


def test_unfold_matches_torch_unfold_4d_and_is_differentiable():
    torch.manual_seed(0)
    x = torch.rand(3, 4, 17, 13, device=DEV, requires_grad=True)
    kw = dict(kernel_size=(3, 2), dilation=2, padding=1, stride=2)
    u = unfoldNd(x, **kw)
    ref = F.unfold(x, **kw)
    assert torch.equal(u, ref)
    g = torch.rand_like(u)
    (gx,) = torch.autograd.grad(u, x, g)
    (gref,) = torch.autograd.grad(ref, x, g)
    close(gx, gref.cpu().numpy(), 1e-6, 1e-6, "fold")
    want = direct.fold(g.cpu().numpy(), 4, (17, 13), **kw)
    close(gx, want, 1e-6, 1e-6, "fold vs oracle")


def test_unfold_host_api_is_bit_exact_and_pipelined():
    """unfoldNd_host: host -> GPU -> host in batch chunks on two streams; same bits as the oracle."""
    from einconv_b200.functionals import unfoldNd_host

    torch.manual_seed(0)
    x = torch.rand(7, 3, 9, 10, 6).pin_memory()
    kw = dict(kernel_size=(3, 2, 3), dilation=(1, 2, 1), padding=(1, 0, 2), stride=(2, 1, 1))
    want = direct.unfold(x.numpy(), kw["kernel_size"], dilation=kw["dilation"], padding=kw["padding"], stride=kw["stride"])
    for chunk in (0, 1, 3, 100):
        got = unfoldNd_host(x, chunk=chunk, **kw)
        assert got.is_pinned() and not got.is_cuda
        assert torch.equal(got, torch.from_numpy(want))
    out = torch.empty_like(got).pin_memory()
    assert unfoldNd_host(x, out=out, **kw) is out and torch.equal(out, got)
    with pytest.raises(ValueError):
        unfoldNd_host(x.to(DEV), **kw)


def _slices_unfold(x, k, s, d, p):
    """Reference-free construction on the GPU: strided slices of the padded input (exact)."""
    N = x.dim() - 2
    pad = []
    for dim in reversed(range(N)):
        pad += [p[dim], p[dim] + s[dim]]  # generous right padding
    xp = F.pad(x, pad)
    out = [1 + (x.shape[2 + i] + 2 * p[i] - (d[i] * (k[i] - 1) + 1)) // s[i] for i in range(N)]
    cols = []
    import itertools

    for taps in itertools.product(*[range(kk) for kk in k]):
        idx = [slice(None), slice(None)]
        for i in range(N):
            start = taps[i] * d[i]
            idx.append(slice(start, start + s[i] * (out[i] - 1) + 1, s[i]))
        cols.append(xp[tuple(idx)].reshape(x.shape[0], x.shape[1], 1, -1))
    return torch.cat(cols, dim=2).reshape(x.shape[0], x.shape[1] * len(cols), -1)


def test_unfold_baseline_c2_full_size_exact(unfold_variant):
    """BASELINE C2 at full size: N=16 C=32 16x56x56 k=3 stride=2 dilation=2 (bit-exact gather)."""
    torch.manual_seed(0)
    x = torch.rand(16, 32, 16, 56, 56).to(DEV)
    got = unfoldNd(x, 3, dilation=2, padding=0, stride=2)
    assert tuple(got.shape) == (16, 864, 4056)
    want = _slices_unfold(x, (3, 3, 3), (2, 2, 2), (2, 2, 2), (0, 0, 0))
    assert torch.equal(got, want)
    # oracle on one sample
    ref = direct.unfold(x[3:4].cpu().numpy(), 3, dilation=2, padding=0, stride=2)
    assert torch.equal(got[3:4].cpu(), torch.from_numpy(ref))


# =============================================================================================
# CUDA graphs: the launch-bound paths are capturable (no hidden synchronisation, no host-side state
# read at replay: tensor maps travel as __grid_constant__ kernel parameters)
# =============================================================================================
def test_cuda_graph_capture_and_replay():
    torch.manual_seed(0)
    x = torch.rand(4, 32, 20, 20, device=DEV)
    w = (torch.rand(48, 32, 3, 3, device=DEV) - 0.5).to(torch.bfloat16)
    wd = torch.rand(32, 1, 3, 3, device=DEV) - 0.5
    v = (torch.rand(4, 48, 20, 20, device=DEV) - 0.5).to(torch.bfloat16)

    def step(xin):
        u = unfoldNd(xin, 3, padding=1)
        y = convNd(xin.to(torch.bfloat16), w, None, padding=1)
        d = convNd(xin, wd, None, padding=1, groups=32)
        gw = evaluator.conv_weight_vjp_raw(xin.to(torch.bfloat16), v, (3, 3), 1, 1, 1, 1)
        k = evaluator.kfc(xin.to(torch.bfloat16), 3, padding=1)          # per-tap TMA kernel (32 tensor maps)
        r = evaluator.kfac_reduce(xin, 3, padding=1)
        return u, y, d, gw, k, r

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):  # warm-up outside capture (cudaFuncSetAttribute, lazy module loading)
            step(x)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()

    static_x = x.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = step(static_x)
    x2 = torch.rand_like(x)
    static_x.copy_(x2)
    graph.replay()
    torch.cuda.synchronize()
    want = step(x2)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], want[0])            # unfold: bit-exact
    for got, ref in zip(outs[1:], want[1:]):        # deterministic kernels: identical bits on replay
        assert torch.equal(got, ref)


# =============================================================================================
# transpose convolutions: unfold (bit-exact) and KFAC-reduce factor
# =============================================================================================
@pytest.mark.parametrize("case", C.TRANSPOSE_UNFOLD_CASES, ids=[c["id"] for c in C.TRANSPOSE_UNFOLD_CASES])
@pytest.mark.parametrize("simplify", SIMPLIFY)
def test_unfold_transpose_bit_exact(case, simplify):
    from einconv_b200.functionals import unfoldNd_transpose

    x = C.draw_input(case)
    want = direct.unfold_transpose(x.numpy(), case["kernel_size"], **case["kwargs"])
    got = unfoldNd_transpose(x.to(DEV), case["kernel_size"], **case["kwargs"], simplify=simplify)
    assert got.dtype == x.dtype and tuple(got.shape) == want.shape
    assert torch.equal(got.cpu(), torch.from_numpy(want))
    golden("unfold_transpose").check(case["id"], got.cpu().numpy(), 0, 0, exact=True)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float64])
def test_unfold_transpose_other_dtypes_and_einsum_hook(dtype):
    from einconv_b200.expressions import conv_transposeNd_unfold

    case = C.TRANSPOSE_UNFOLD_CASES[5]
    x = C.draw_input(case).to(dtype)
    raw = x.view(torch.int16).numpy() if dtype == torch.bfloat16 else x.numpy()
    want = direct.unfold_transpose(raw, case["kernel_size"], **case["kwargs"])
    eq, ops, shape = conv_transposeNd_unfold.einsum_expression(x.to(DEV), case["kernel_size"], **case["kwargs"])
    got = torch.einsum(eq, *ops).reshape(shape).cpu()
    got_raw = got.view(torch.int16).numpy() if dtype == torch.bfloat16 else got.numpy()
    assert np.array_equal(got_raw, want)


@pytest.mark.parametrize("case", C.TRANSPOSE_FACTOR_CASES, ids=[c["id"] for c in C.TRANSPOSE_FACTOR_CASES])
@pytest.mark.parametrize("simplify", SIMPLIFY)
def test_kfac_reduce_transpose_fp32(case, simplify):
    """Through the reference's calling convention: einsum_expression + torch.einsum(...).reshape(shape);
    the reference's tolerance for factors (rtol 5e-5)."""
    from einconv_b200.expressions import conv_transposeNd_kfac_reduce

    x = C.draw_input(case)
    want = direct.kfac_reduce_transpose(x.numpy(), case["kernel_size"], **case["kwargs"])
    eq, ops, shape = conv_transposeNd_kfac_reduce.einsum_expression(
        x.to(DEV), case["kernel_size"], **case["kwargs"], simplify=simplify
    )
    got = torch.einsum(eq, *ops).reshape(shape)
    assert tuple(got.shape) == want.shape
    close(got, want, 5e-5, 1e-7, "kfac_reduce_transpose")
    golden("kfac_reduce_transpose").check(case["id"], got.cpu().numpy(), 5e-5, 1e-7)


def test_kfac_reduce_transpose_tensor_core_stage2():
    """A >= 256: window sums + TF32 Gram on the TMA kernel (1e-3 max|ref|)."""
    torch.manual_seed(0)
    x = torch.rand(40, 64, 9, 9)
    kw = dict(stride=2, padding=1, output_padding=1)
    want = direct.kfac_reduce_transpose(x.numpy(), 3, **kw)
    evaluator.set_fp32_math("tf32")
    try:
        got = evaluator.kfac_reduce_transpose(x.to(DEV), 3, **kw)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, 1e-3, "kfac_reduce_transpose tf32")


# =============================================================================================
# forward
# =============================================================================================
FWD = C.CONV_CASES + C.SPECIAL_CASES + C.EDGE_CASES


@pytest.mark.parametrize("case", FWD, ids=[c["id"] for c in FWD])
@pytest.mark.parametrize("simplify", SIMPLIFY)
def test_forward_fp32(case, simplify):
    """fp32, EINCONV_MATH_FP32: rtol 1e-5 (+ atol 1e-5 for cancellation-free rand inputs)."""
    x, w, b = C.draw_forward(case)
    want = direct.conv_forward(x.numpy(), w.numpy(), None if b is None else b.numpy(), **case["kwargs"])
    got = convNd(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV), **case["kwargs"], simplify=simplify)
    assert got.dtype == torch.float32
    close(got, want, 1e-5, 1e-5, case["id"])
    golden("forward").check(case["id"], got.cpu().numpy(), 1e-5, 1e-5)


@pytest.mark.parametrize("case", [C.CONV_CASES[9], C.CONV_CASES[18], C.SPECIAL_CASES[5], C.EDGE_CASES[6]],
                         ids=lambda c: c["id"])
def test_forward_fp64(case):
    x, w, b = C.draw_forward(case)
    x, w = x.double(), w.double()
    b = None if b is None else b.double()
    want = direct.conv_forward(x.numpy(), w.numpy(), None if b is None else b.numpy(), **case["kwargs"])
    got = convNd(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV), **case["kwargs"])
    assert got.dtype == torch.float64
    close(got, want, 1e-10, 1e-12, case["id"])


@pytest.mark.parametrize("dtype,ulp", [(torch.bfloat16, 2.0**-8), (torch.float16, 2.0**-11)])
@pytest.mark.parametrize("case", [C.CONV_CASES[8], C.CONV_CASES[11], C.SPECIAL_CASES[5], C.SPECIAL_CASES[0]],
                         ids=lambda c: c["id"])
@pytest.mark.parametrize("simplify", SIMPLIFY)
def test_forward_16bit(case, dtype, ulp, simplify):
    """16-bit I/O, fp32 accumulation: one output rounding against the fp64 oracle on the same bits."""
    x, w, b = C.draw_forward(case)
    x, w = x.to(dtype), w.to(dtype)
    b = None if b is None else b.to(dtype)
    want = direct.conv_forward(x.double().numpy(), w.double().numpy(),
                               None if b is None else b.double().numpy(), **case["kwargs"])
    got = convNd(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV), **case["kwargs"], simplify=simplify)
    assert got.dtype == dtype
    close(got, want, ulp, 1e-4, case["id"])


def test_forward_matches_torch_conv_and_module():
    """Mirror of test/modules/test_conv.py: nn.ConvNd -> ConvNd.from_nn_Conv, atol 1e-6 rtol 1e-5."""
    torch.manual_seed(0)
    for ref in (
        torch.nn.Conv1d(6, 9, 5, stride=2, padding=1, dilation=2, groups=3),
        torch.nn.Conv2d(6, 9, (5, 4), padding="same", dilation=(1, 2), groups=3),
        torch.nn.Conv3d(3, 4, 3, stride=2, padding=1, bias=False),
    ):
        ref = ref.to(DEV)
        N = ref.weight.dim() - 2
        x = torch.rand(2, ref.in_channels, *([13, 11, 9][:N]), device=DEV)
        for simplify in SIMPLIFY:
            mine = ConvNd.from_nn_Conv(ref, simplify=simplify)
            with torch.backends.cudnn.flags(allow_tf32=False):
                want = ref(x)
            close(mine(x), want.detach().cpu().numpy(), 1e-5, 1e-5, f"module N={N}")


def test_forward_baseline_c1_shape_and_linearity():
    """BASELINE C1 (N=32 C=64 56x56 k3 p1 fp32) at full size: against cuDNN fp32 and linearity."""
    torch.manual_seed(0)
    x = torch.rand(32, 64, 56, 56, device=DEV)
    w = torch.rand(64, 64, 3, 3, device=DEV)
    y = convNd(x, w, padding=1)
    with torch.backends.cudnn.flags(allow_tf32=False):
        ref = F.conv2d(x.double(), w.double(), padding=1)
    close(y, ref.cpu().numpy(), 2e-5, 1e-4, "C1 vs fp64 conv")
    y2 = convNd(2.0 * x, w, padding=1)
    assert torch.equal(y2, 2.0 * y)  # scaling by a power of two is exact
    # oracle on a slice of the batch
    want = direct.conv_forward(x[:1].cpu().numpy(), w.cpu().numpy(), None, padding=1)
    close(y[:1], want, 1e-5, 1e-4, "C1 slice vs oracle")


# =============================================================================================
# VJPs through the reference's calling convention: torch.einsum(equation, *operands).reshape(shape)
# =============================================================================================
VJP = C.VJP_CASES + C.SPECIAL_CASES


@pytest.mark.parametrize("case", VJP, ids=[c["id"] for c in VJP])
@pytest.mark.parametrize("simplify", SIMPLIFY)
def test_input_vjp_fp32(case, simplify):
    """Reference tolerance for the input VJP: atol 5e-7 (test_convNd_input_vjp.py:55) plus rtol 1e-5."""
    w, x, v = C.draw_vjp(case, forward_out_shape(case))
    want = direct.conv_input_vjp(w.numpy(), v.numpy(), tuple(x.shape[2:]), **case["kwargs"])
    eq, ops, shape = convNd_input_vjp.einsum_expression(
        w.to(DEV), v.to(DEV), tuple(x.shape[2:]), **case["kwargs"], simplify=simplify
    )
    before = nat.launch_count()
    got = torch.einsum(eq, *ops).reshape(shape)
    assert nat.launch_count() > before, "einsum did not reach the native kernels"
    close(got, want, 1e-5, 5e-6, case["id"])
    if golden("input_vjp").has(case["id"]):
        golden("input_vjp").check(case["id"], got.cpu().numpy(), 1e-5, 5e-6)


@pytest.mark.parametrize("case", VJP, ids=[c["id"] for c in VJP])
@pytest.mark.parametrize("simplify", SIMPLIFY)
def test_weight_vjp_fp32(case, simplify):
    """fp32 weight VJP: rtol 2e-5 (reductions over up to 10^5 terms, SURVEY.md 7 'precision')."""
    w, x, v = C.draw_vjp(case, forward_out_shape(case))
    want = direct.conv_weight_vjp(x.numpy(), v.numpy(), tuple(w.shape[2:]), **case["kwargs"])
    eq, ops, shape = convNd_weight_vjp.einsum_expression(
        x.to(DEV), v.to(DEV), tuple(w.shape[2:]), **case["kwargs"], simplify=simplify
    )
    before = nat.launch_count()
    got = torch.einsum(eq, *ops).reshape(shape)
    assert nat.launch_count() > before
    assert tuple(got.shape) == tuple(w.shape)
    close(got, want, 2e-5, 1e-5, case["id"])
    if golden("weight_vjp").has(case["id"]):
        golden("weight_vjp").check(case["id"], got.cpu().numpy(), 2e-5, 1e-5)


def test_einsum_result_has_the_shape_torch_would_return():
    x = torch.rand(2, 4, 9, 9, device=DEV)
    w = torch.rand(6, 2, 3, 3, device=DEV)
    eq, ops, shape = convNd_forward.einsum_expression(x, w, groups=2, simplify=False)
    out = torch.einsum(eq, *ops)
    assert tuple(out.shape) == (2, 2, 3, 7, 7)  # n g c_out o0 o1
    assert tuple(out.reshape(shape).shape) == (2, 6, 7, 7)
    # a user-composed network with a pattern still evaluates (written-out pattern, torch on CUDA)
    pat = einconv_b200.index_pattern(9, 3, device=x.device, dtype=x.dtype)
    manual = torch.einsum("nchw,koh->nckow", x, pat)
    assert tuple(manual.shape) == (2, 4, 3, 7, 9)


def test_autograd_through_convnd_matches_torch():
    torch.manual_seed(0)
    for kw in (dict(padding=1), dict(stride=2, padding=2, dilation=2, groups=2),
               dict(padding=1, groups=8)):
        groups = kw.get("groups", 1)
        x = torch.rand(3, 8, 12, 10, device=DEV, requires_grad=True)
        w = torch.rand(16, 8 // groups, 3, 3, device=DEV, requires_grad=True)
        b = torch.rand(16, device=DEV, requires_grad=True)
        for simplify in SIMPLIFY:
            y = convNd(x, w, b, **kw, simplify=simplify)
            g = torch.rand_like(y)
            gx, gw, gb = torch.autograd.grad(y, (x, w, b), g)
            with torch.backends.cudnn.flags(allow_tf32=False):
                yr = F.conv2d(x.double(), w.double(), b.double(), **kw)
                rx, rw, rb = torch.autograd.grad(yr, (x, w, b), g.double())
            close(gx, rx.cpu().numpy(), 1e-5, 1e-5, f"gx {kw}")
            close(gw, rw.cpu().numpy(), 2e-5, 1e-4, f"gw {kw}")
            close(gb, rb.cpu().numpy(), 1e-5, 1e-4, f"gb {kw}")


def test_weight_vjp_baseline_c4_reduced_bf16():
    """BASELINE C4 geometry (3-d, k3 p1) at reduced size in bf16: fp32 accumulate, bf16 output."""
    torch.manual_seed(0)
    x = torch.rand(2, 16, 6, 28, 28).bfloat16()
    v = torch.rand(2, 16, 6, 28, 28).bfloat16()
    want = direct.conv_weight_vjp(x.double().numpy(), v.double().numpy(), 3, padding=1)
    for simplify in SIMPLIFY:
        eq, ops, shape = convNd_weight_vjp.einsum_expression(x.to(DEV), v.to(DEV), 3, padding=1, simplify=simplify)
        got = torch.einsum(eq, *ops).reshape(shape)
        assert got.dtype == torch.bfloat16
        close(got, want, 2.0**-8, 1e-2, "C4 reduced")


# =============================================================================================
# Kronecker factors
# =============================================================================================
def _factor_truth(x, kernel_size, kwargs, kind):
    """Ground truth of the reference's tests (test_convNd_kfc.py:47-52, test_convNd_kfac_reduce.py:48-53):
    unfold (already proven bit-exact above), then the outer products in fp64 on the GPU."""
    groups = kwargs.get("groups", 1)
    ukw = {k: v for k, v in kwargs.items() if k != "groups"}
    u = unfoldNd(x.to(DEV), kernel_size, **ukw).double()
    B = u.shape[0]
    u = u.reshape(B, groups, u.shape[1] // groups, u.shape[2])
    if kind == "kfc":
        return torch.einsum("ngik,ngjk->gij", u, u) / B
    m = u.mean(-1)
    return torch.einsum("ngi,ngj->gij", m, m) / B


@pytest.mark.parametrize("case", C.FACTOR_CASES, ids=[c["id"] for c in C.FACTOR_CASES])
@pytest.mark.parametrize("simplify", SIMPLIFY)
@pytest.mark.parametrize("kind", ["kfc", "kfac_reduce"])
def test_factors_fp32(case, simplify, kind):
    """rtol 5e-5 as in the reference (test_convNd_kfc.py:59, test_convNd_kfac_reduce.py:60)."""
    x = C.draw_input(case)
    mod = convNd_kfc if kind == "kfc" else convNd_kfac_reduce
    eq, ops, shape = mod.einsum_expression(x.to(DEV), case["kernel_size"], **case["kwargs"], simplify=simplify)
    before = nat.launch_count()
    got = torch.einsum(eq, *ops).reshape(shape)
    assert nat.launch_count() > before
    truth = _factor_truth(x, case["kernel_size"], case["kwargs"], kind)
    close(got, truth.cpu().numpy(), 5e-5, 1e-6 if kind == "kfc" else 1e-8, f"{kind} {case['id']}")
    if golden(kind).has(case["id"]):
        golden(kind).check(case["id"], got.cpu().numpy(), 5e-5, 1e-6)
    if len(case["x"]) <= 4:
        fn = direct.kfc if kind == "kfc" else direct.kfac_reduce
        close(got, fn(x.numpy(), case["kernel_size"], **case["kwargs"]), 5e-5, 1e-6, f"{kind} oracle")


def test_factors_resnet_layers_reduced_batch():
    """A few ResNet-50 geometries of BASELINE C5 at batch 4."""
    layers = [(3, 56, 7, 2, 3), (64, 28, 3, 1, 1), (128, 28, 3, 2, 1), (256, 14, 1, 1, 0), (256, 14, 1, 2, 0)]
    for n, (c, h, k, s, p) in enumerate(layers):
        torch.manual_seed(n)
        x = torch.rand(4, c, h, h)
        kw = dict(stride=s, padding=p)
        for kind, fn in (("kfc", evaluator.kfc), ("kfac_reduce", evaluator.kfac_reduce)):
            got = fn(x.to(DEV), k, **kw)
            truth = _factor_truth(x, k, kw, kind)
            close(got, truth.cpu().numpy(), 5e-5, 1e-6 if kind == "kfc" else 1e-9, f"{kind} layer {n}")
            assert torch.allclose(got, got.transpose(1, 2), rtol=1e-6, atol=1e-7)  # symmetric


def test_factor_sweep_graphs_match_eager_and_track_input_updates():
    """FactorSweep (in-place bucket outputs, one CUDA graph per bucket) reproduces sharded_factors bit
    for bit, and a replay sees new contents of the (static) input tensors."""
    from einconv_b200 import distributed as D

    torch.manual_seed(0)
    xs = [torch.rand(4, 16, 12, 12, device=DEV), torch.rand(4, 32, 6, 6, device=DEV), torch.rand(4, 3, 20, 20, device=DEV)]
    specs = [D.LayerSpec(3, padding=1), D.LayerSpec(1), D.LayerSpec(5, stride=2, padding=2)]
    sweep = D.FactorSweep(xs, specs, global_batch=4, bucket_bytes=1 << 16)
    assert sweep._graphs is not None and len(sweep.buckets) >= 2 and sweep.launches_per_run > 0
    for trial in range(2):
        if trial:
            for x in xs:
                x.copy_(torch.rand_like(x))
        got = sweep.run()
        want = D.sharded_factors(xs, specs, global_batch=4)
        for kind in ("kfc", "kfac_reduce"):
            for a, b in zip(got[kind], want[kind]):
                assert torch.equal(a, b)


def test_factor_sharding_identity_on_one_gpu():
    """Sum of shard factors computed with the global scale == the full-batch factor."""
    torch.manual_seed(0)
    x = torch.rand(8, 6, 14, 14, device=DEV)
    full_k = evaluator.kfc(x, 3, padding=1)
    full_r = evaluator.kfac_reduce(x, 3, padding=1)
    parts_k = sum(evaluator.kfc(x[i : i + 2], 3, padding=1, scale=1.0 / 8) for i in range(0, 8, 2))
    parts_r = sum(evaluator.kfac_reduce(x[i : i + 2], 3, padding=1, scale=1.0 / (8 * 196**2)) for i in range(0, 8, 2))
    close(parts_k, full_k.cpu().numpy(), 1e-5, 1e-6, "kfc shards")
    close(parts_r, full_r.cpu().numpy(), 1e-5, 1e-9, "kfac_reduce shards")


# =============================================================================================
# tensor-core (tcgen05 / TMEM) paths: bf16 / fp16 natively, fp32 via TF32 when opted in
# =============================================================================================
TC_CASES = [
    # (x shape, w shape, kwargs)
    ((2, 32, 20, 20), (48, 32, 3, 3), dict(padding=1)),
    ((3, 16, 23, 17), (40, 16, 3, 3), dict(padding=1, stride=2)),
    ((2, 24, 19, 19), (16, 12, 5, 3), dict(padding=(2, 0), dilation=(1, 2), groups=2)),
    ((2, 8, 300), (24, 8, 7), dict(padding=3, stride=2)),
    ((1, 8, 9, 12, 10), (16, 8, 3, 3, 3), dict(padding=1)),
    ((2, 70, 9, 9), (130, 70, 1, 1), dict()),
    ((1, 64, 14, 14), (64, 64, 3, 3), dict(padding=1)),
]
TC_IDS = ["x".join(map(str, c[0])) + "-w" + "x".join(map(str, c[1])) for c in TC_CASES]


def _tc_inputs(case, dtype):
    xs, ws, kw = case
    torch.manual_seed(0)
    x = (torch.rand(*xs) - 0.3).to(dtype)
    w = (torch.rand(*ws) - 0.5).to(dtype)
    b = (torch.rand(ws[0]) - 0.5).to(dtype)
    return x, w, b, kw


def _rel_to_max(got, want, tol, what):
    got = got.detach().double().cpu().numpy()
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.abs(got - want).max()
    ref = np.abs(want).max()
    assert err <= tol * ref, f"{what}: max|d| = {err:.3e}, max|ref| = {ref:.3e}, ratio {err / ref:.3e} > {tol}"


@pytest.mark.parametrize("case", TC_CASES, ids=TC_IDS)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_tc_forward_16bit(case, dtype):
    """kind::f16 MMA, fp32 accumulate in TMEM, one rounding to the 16-bit output:
    max|d| <= 2^-7 (bf16) / 2^-10 (fp16) of max|ref| against the fp64 oracle on the same bits."""
    x, w, b, kw = _tc_inputs(case, dtype)
    g = nat.make_geom(1, 1, 1, 1, (1,), (1,), (1,), (1,), (0,), (1,))
    want = direct.conv_forward(x.double().numpy(), w.double().numpy(), b.double().numpy(), **kw)
    got = convNd(x.to(DEV), w.to(DEV), b.to(DEV), **kw)
    _rel_to_max(got, want, 2.0**-7 if dtype == torch.bfloat16 else 2.0**-10, "tc fwd")


@pytest.mark.parametrize("case", TC_CASES, ids=TC_IDS)
def test_tc_forward_tf32(case):
    """kind::tf32 (operands rounded to 10-bit mantissa, fp32 accumulate): max|d| <= 1e-3 max|ref|
    (SURVEY.md 8c).  Opt-in through set_fp32_math / torch.backends.cuda.matmul.allow_tf32."""
    x, w, b, kw = _tc_inputs(case, torch.float32)
    want = direct.conv_forward(x.numpy(), w.numpy(), b.numpy(), **kw)
    evaluator.set_fp32_math("tf32")
    try:
        got = convNd(x.to(DEV), w.to(DEV), b.to(DEV), **kw)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, 1e-3, "tc fwd tf32")
    exact = convNd(x.to(DEV), w.to(DEV), b.to(DEV), **kw)  # default stays exact fp32
    close(exact, want, 1e-5, 1e-5, "fp32 default")


@pytest.mark.parametrize("case", TC_CASES, ids=TC_IDS)
@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_tc_weight_vjp(case, mode):
    xs, ws, kw = case
    torch.manual_seed(1)
    dtype = torch.bfloat16 if mode == "bf16" else torch.float32
    x = (torch.rand(*xs) - 0.3).to(dtype)
    yshape = forward_out_shape(dict(x=xs, w=ws, kwargs=kw))
    v = (torch.rand(*yshape) - 0.5).to(dtype)
    want = direct.conv_weight_vjp(x.double().numpy(), v.double().numpy(), tuple(ws[2:]), **kw)
    evaluator.set_fp32_math("tf32" if mode == "tf32" else "auto")
    try:
        eq, ops, shape = convNd_weight_vjp.einsum_expression(x.to(DEV), v.to(DEV), tuple(ws[2:]), **kw)
        got = torch.einsum(eq, *ops).reshape(shape)
    finally:
        evaluator.set_fp32_math("auto")
    assert tuple(got.shape) == tuple(ws)
    _rel_to_max(got, want, 2.0**-7 if mode == "bf16" else 1e-3, f"tc wgrad {mode}")


@pytest.mark.parametrize("case", TC_CASES, ids=TC_IDS)
@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_tc_kfc(case, mode):
    xs, ws, kw = case
    torch.manual_seed(2)
    dtype = torch.bfloat16 if mode == "bf16" else torch.float32
    x = torch.rand(*xs).to(dtype)
    want = direct.kfc(x.double().numpy(), tuple(ws[2:]), **kw)
    evaluator.set_fp32_math("tf32" if mode == "tf32" else "auto")
    try:
        got = evaluator.kfc(x.to(DEV), tuple(ws[2:]), **kw)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, 2.0**-7 if mode == "bf16" else 1e-3, f"tc kfc {mode}")


@pytest.mark.parametrize("variant", ["default", "EINCONV_KFC_FINISH_BC=16", "EINCONV_KFC_FINISH_TILE=1",
                                     "EINCONV_KFC_FUSED_CONVERT=1"])
@pytest.mark.parametrize("shape,stride", [((2, 256, 7, 7), 1), ((3, 272, 9, 8), 1), ((2, 256, 14, 14), 2)])
def test_kfc_finish_and_copy_variants_many_channels(variant, shape, stride, monkeypatch):
    """Factors with many channels and several taps: the output-space finish kernel (blocks of 8 or 16 channels,
    all taps gathered in shared memory, whole factor rows and their mirror written) against the tile-space
    kernel, and the converting flat copy against scale_to_half + copy; every variant against the oracle
    (1e-3 max|ref|: fp32 tensors in TF32 mode).  272 channels: a ragged last channel block."""
    torch.manual_seed(3)
    x = torch.rand(*shape) - 0.4
    want = direct.kfc(x.double().numpy(), (3, 3), padding=1, stride=stride)
    if variant != "default":
        k, v = variant.split("=")
        monkeypatch.setenv(k, v)
    evaluator.clear_plan_cache()
    evaluator.set_fp32_math("tf32")
    try:
        got = evaluator.kfc(x.to(DEV), (3, 3), padding=1, stride=stride)
    finally:
        evaluator.set_fp32_math("auto")
        evaluator.clear_plan_cache()
    _rel_to_max(got, want, 1e-3, f"kfc {variant} {shape}")
    g = got.double().cpu().numpy()
    assert np.array_equal(g, np.swapaxes(g, -1, -2)), "factor is not exactly symmetric"


@pytest.mark.parametrize("magnitude", [1.0, 3.0e4, 2.0e-7])
def test_kfc_scaled_fp16_operands_match_tf32_accuracy(magnitude, monkeypatch):
    """fp32 KFC in TF32 mode runs on fp16 copies scaled by a power of two (csrc/convert.cu): same
    tolerance as kind::tf32 (1e-3 max|ref|) for any magnitude of the data, and not worse than the
    kind::tf32 route (EINCONV_KFC_TF32=1) by more than a rounding."""
    torch.manual_seed(0)
    x = ((torch.rand(3, 32, 18, 18) - 0.3) * magnitude)
    want = direct.kfc(x.double().numpy(), (3, 3), padding=1)
    evaluator.set_fp32_math("tf32")
    try:
        got16 = evaluator.kfc(x.to(DEV), (3, 3), padding=1)
        monkeypatch.setenv("EINCONV_KFC_TF32", "1")
        evaluator.clear_plan_cache()  # plans remember the workspace of the route chosen when they were made
        got32 = evaluator.kfc(x.to(DEV), (3, 3), padding=1)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got16, want, 1e-3, "kfc scaled fp16")
    _rel_to_max(got32, want, 1e-3, "kfc tf32")
    ref = np.abs(want).max()
    e16 = np.abs(got16.double().cpu().numpy() - want).max() / ref
    e32 = np.abs(got32.double().cpu().numpy() - want).max() / ref
    assert e16 <= 2.0 * e32 + 1e-6, (e16, e32)


def test_tc_kernels_are_selected():
    g = nat.make_geom(2, 1, 32, 48, (20, 20), (3, 3), (1, 1), (1, 1), (1, 1), (20, 20))
    lib = nat.lib()
    # stride 1: TMA-fed implicit GEMM over the padded channels-last copy (forward and input VJP)
    assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, g, 1, nat.MATH_AUTO, 0) == b"tcgen05_tma_conv_fwd"
    # fp32 in TF32 mode, 2-d stride 1: the channels-first row kernel (csrc/conv_rows.cu, no packed copy);
    # EINCONV_NO_CONV_ROWS=1 falls back to the TMA kernel over the channels-last copy
    assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, g, 0, nat.MATH_TF32, 0) == b"tcgen05_rows_conv_fwd"
    assert lib.einconv_selected_kernel(nat.OP_CONV_INPUT_VJP, g, 1, nat.MATH_AUTO, 0) == b"tcgen05_tma_conv_input_vjp"
    assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, g, 0, nat.MATH_FP32, 0) == b"ffma_gather_gemm"
    assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, g, 1, nat.MATH_AUTO, 1) == b"ffma_gather_gemm"
    # stride 2 forward: the same TMA kernel over phase sub-images of the source (csrc/conv_tma.cu, Role::strided);
    # the software-gather tcgen05 kernel is only the fallback (EINCONV_CTMA_NO_STRIDE=1)
    gs = nat.make_geom(2, 1, 32, 48, (20, 20), (3, 3), (2, 2), (1, 1), (1, 1), (10, 10))
    assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, gs, 1, nat.MATH_AUTO, 0) == b"tcgen05_tma_conv_fwd"
    assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, gs, 0, nat.MATH_TF32, 0) == b"tcgen05_tma_conv_fwd"
    import os
    os.environ["EINCONV_CTMA_NO_STRIDE"] = "1"
    try:
        assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, gs, 1, nat.MATH_AUTO, 0) == b"tcgen05_fwd_f16"
        assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, gs, 0, nat.MATH_TF32, 0) == b"tcgen05_fwd_tf32"
    finally:
        del os.environ["EINCONV_CTMA_NO_STRIDE"]
    # reduction networks: TMA-fed kernel (aligned shifted copies make any row length usable)
    assert lib.einconv_selected_kernel(nat.OP_KFC, g, 1, nat.MATH_AUTO, 0) == b"tcgen05_tma_kfc"
    # fp32 in TF32 mode: scaled fp16 operand copies (same 10-bit mantissa) for kernels with several taps
    assert lib.einconv_selected_kernel(nat.OP_KFC, g, 0, nat.MATH_TF32, 0) == b"f16scaled+tcgen05_tma_kfc"
    g11 = nat.make_geom(2, 1, 64, 48, (20, 20), (1, 1), (1, 1), (1, 1), (0, 0), (20, 20))
    assert lib.einconv_selected_kernel(nat.OP_KFC, g11, 0, nat.MATH_TF32, 0) == b"tcgen05_tma_kfc"
    assert lib.einconv_selected_kernel(nat.OP_KFC, g, 0, nat.MATH_FP32, 0) == b"ffma_gather_gemm"
    g24 = nat.make_geom(2, 1, 32, 48, (24, 24), (3, 3), (1, 1), (1, 1), (1, 1), (24, 24))
    assert lib.einconv_selected_kernel(nat.OP_KFC, g24, 1, nat.MATH_AUTO, 0) == b"tcgen05_tma_kfc"
    # stride 1: slab kernel on channels-last padded copies (fp32 in TF32 mode as scaled fp16); stride > 1: per-tap TMA tiles
    assert lib.einconv_selected_kernel(nat.OP_CONV_WEIGHT_VJP, g24, 1, nat.MATH_AUTO, 0) == b"tcgen05_slab_wgrad"
    assert lib.einconv_selected_kernel(nat.OP_CONV_WEIGHT_VJP, g24, 0, nat.MATH_TF32, 0) == b"tcgen05_slab_wgrad"
    g24s = nat.make_geom(2, 1, 32, 48, (24, 24), (3, 3), (2, 2), (1, 1), (1, 1), (12, 12))
    assert lib.einconv_selected_kernel(nat.OP_CONV_WEIGHT_VJP, g24s, 1, nat.MATH_AUTO, 0) == b"tcgen05_tma_wgrad"
    assert lib.einconv_selected_kernel(nat.OP_CONV_WEIGHT_VJP, g24, 1, nat.MATH_AUTO, 1) == b"ffma_gather_gemm"


# ---- depthwise stencils (csrc/depthwise.cu): warp-per-strip kernel and the per-column kernel --------
DW_CASES = [
    ((2, 256, 28, 28), 3, dict(padding=1)),          # C3 shape (one strip, two row blocks)
    ((2, 8, 40, 75), 3, dict(padding=1)),            # three column strips, ragged last strip
    ((3, 6, 19, 33), 5, dict(padding=2)),            # 5x5
    ((2, 4, 17, 30), 7, dict(padding=3)),            # 7x7
    ((2, 5, 1, 50), (1, 3), dict(padding=(0, 1))),   # 1x3
    ((2, 6, 20, 20), 3, dict(padding=0)),            # no padding: output smaller than input
    ((2, 6, 9, 9), 3, dict(padding=2)),              # padding = K - 1
    ((2, 6, 21, 22), 3, dict(padding=1, stride=2)),  # stride 2: per-column kernel
    # rows a multiple of 4 wide with "same" padding along W: the 16-byte vectorised kernel (fp32)
    ((2, 6, 20, 36), 3, dict(padding=1)),            # 9 lanes per row, 3 units per warp
    ((2, 3, 33, 128), 3, dict(padding=1)),           # 32 lanes per row, three row blocks
    ((3, 5, 13, 8), 5, dict(padding=2)),             # 2 lanes per row, 16 units per warp
    ((2, 4, 30, 24), 7, dict(padding=3)),            # 7x7
    ((2, 6, 10, 12), 3, dict(padding=(2, 1))),       # taller output than input
    ((2, 6, 10, 12), 3, dict(padding=(0, 1))),       # no padding along H
    ((1, 3, 5, 4), 3, dict(padding=1)),              # a single lane per row
]


@pytest.mark.parametrize("case", DW_CASES, ids=lambda c: "x".join(map(str, c[0])) + f"-k{c[1]}")
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("variant", ["auto", "warp", "column"])
def test_depthwise_kernels(case, dtype, variant, monkeypatch):
    """fp32: CUDA-core FMA in a fixed order, rtol 1e-5 like the reference's tests; bf16: 2^-7 max|ref|."""
    if variant == "column":
        monkeypatch.setenv("EINCONV_DW_DIRECT", "1")
    elif variant == "warp":
        monkeypatch.setenv("EINCONV_DW_NO_VEC", "1")
    xs, k, kw = case
    C = xs[1]
    ks = (k, k) if isinstance(k, int) else k
    torch.manual_seed(0)
    x = (torch.rand(*xs) - 0.3).to(dtype)
    w = (torch.rand(C, 1, *ks) - 0.5).to(dtype)
    b = (torch.rand(C) - 0.5).to(dtype)
    want = direct.conv_forward(x.double().numpy(), w.double().numpy(), b.double().numpy(), groups=C, **kw)
    got = convNd(x.to(DEV), w.to(DEV), b.to(DEV), groups=C, **kw)
    torch.manual_seed(1)
    v = (torch.rand(*want.shape) - 0.5).to(dtype)
    want_gx = direct.conv_input_vjp(w.double().numpy(), v.double().numpy(), tuple(xs[2:]), groups=C, **kw)
    gx = evaluator.conv_input_vjp_raw(
        w.to(DEV), v.to(DEV), tuple(xs[2:]), kw.get("stride", 1), kw.get("padding", 0), 1, C
    )
    if dtype == torch.float32:
        close(got, want, 1e-5, 1e-6, "depthwise fwd")
        close(gx, want_gx, 1e-5, 1e-6, "depthwise input vjp")
    else:
        _rel_to_max(got, want, 2.0**-7, "depthwise fwd bf16")
        _rel_to_max(gx, want_gx, 2.0**-7, "depthwise input vjp bf16")


# ---- TMA-fed implicit GEMM for forward / input VJP (csrc/conv_tma.cu): stride 1, any N --------------
CTMA_CASES = [
    ((2, 32, 20, 20), (48, 32, 3, 3), dict(padding=1)),
    ((2, 3, 30, 31), (20, 3, 7, 7), dict(padding=3)),                        # C_in = 3: 32-byte slices, 49 taps
    ((2, 24, 19, 19), (16, 12, 5, 3), dict(padding=(2, 0), dilation=(1, 2), groups=2)),
    ((3, 16, 50), (24, 16, 5), dict(padding=2)),                             # N = 1
    ((2, 16, 9, 10, 11), (40, 16, 3, 3, 3), dict(padding=2, dilation=2)),    # N = 3, dilated
    ((1, 8, 5, 6, 7, 6), (16, 8, 2, 3, 2, 3), dict(padding=1)),              # N = 4 (tap groups: level < N)
    ((2, 300, 7, 7), (520, 300, 3, 3), dict(padding=1)),                     # 3 n-tiles, 5 channel slices, streaming
    ((2, 16, 9, 8), (16, 16, 4, 4), dict(padding="same")),                   # asymmetric padding
    ((2, 8, 12, 12), (8, 8, 3, 3), dict(padding=0)),
    ((1, 64, 56, 56), (64, 64, 3, 3), dict(padding=1)),                      # C1 shape, one sample
    ((2, 64, 14, 14), (256, 64, 1, 1), dict()),                              # 1x1, BN = 256
    # 49..64 output channels per group: the weight VJP runs four taps per MMA ("quad" mode)
    ((2, 32, 10, 11), (64, 32, 3, 3), dict(padding=1)),
    ((2, 16, 6, 7, 8), (64, 16, 3, 3, 3), dict(padding=1)),                  # tap groups per depth plane
    ((2, 16, 12, 13), (56, 16, 2, 4), dict(padding=1)),                      # even kernel width, 56 of 64 columns
    ((2, 16, 11, 12), (64, 16, 3, 5), dict(padding=(1, 6), dilation=(1, 3))),  # dilated pairs, groups per row
    ((3, 16, 40), (64, 16, 3), dict(padding=1)),                             # N = 1
    ((2, 32, 9, 9), (128, 16, 3, 3), dict(padding=1, groups=2)),             # two conv groups of 64
    # innermost extents that are multiples of 8: the weight VJP fills its slabs from the channels-first
    # tensors with producer warps (no packed copies); 3 / 5 / 7 taps per kernel row: N = 192 / 256 MMAs
    ((2, 64, 4, 8, 16), (64, 64, 3, 3, 3), dict(padding=1)),                 # C4 in small
    ((2, 32, 10, 16), (128, 32, 3, 3), dict(padding=1)),                     # two cotangent blocks per tile
    ((2, 24, 12, 24), (40, 24, 3, 3), dict(padding=(1, 2), dilation=(1, 2))),  # BN = 48, dilated
    ((2, 16, 9, 16), (128, 8, 5, 5), dict(padding=2, groups=2)),             # 5 taps per row: units of 3 + 2
    ((1, 8, 10, 24), (64, 8, 1, 7), dict(padding=(0, 3))),                   # 7 taps per row: units of 4 + 3
    ((3, 72, 5, 8), (64, 72, 3, 3), dict(padding=0)),                        # two channel blocks, no padding
]
CTMA_IDS = ["x".join(map(str, c[0])) + "-w" + "x".join(map(str, c[1])) for c in CTMA_CASES]


@pytest.fixture(params=["auto", "stream", "fused", "fused-stream"])
def ctma_weights(request, monkeypatch):
    """Weights resident in shared memory when they fit ("auto") or always streamed per stage; "fused": the
    A slabs are filled from the channels-first source by producer warps instead of TMA boxes of a packed
    copy (opt-in route; applies where the innermost source rows are multiples of 16 bytes)."""
    if "stream" in request.param:
        monkeypatch.setenv("EINCONV_CTMA_STREAM", "1")
    else:
        monkeypatch.delenv("EINCONV_CTMA_STREAM", raising=False)
    if "fused" in request.param:
        monkeypatch.setenv("EINCONV_CTMA_FUSED", "1")
    else:
        monkeypatch.delenv("EINCONV_CTMA_FUSED", raising=False)
    return request.param


@pytest.mark.parametrize("case", CTMA_CASES, ids=CTMA_IDS)
@pytest.mark.parametrize("mode", ["bf16", "fp16", "tf32", "tf32-native", "tf32-half"])
def test_conv_tma_forward_and_input_vjp(case, mode, ctma_weights, monkeypatch):
    """Same tolerances as the other tensor-core tests: against the fp64 oracle on the same rounded
    inputs, max|d| <= 2^-7 (bf16) / 2^-10 (fp16) / 1e-3 (TF32) of max|ref|.  fp32 tensors in TF32 mode:
    "tf32" = the dispatcher's choice, "tf32-native" = kind::tf32 MMAs, "tf32-half" = power-of-two-scaled
    fp16 operand copies (same 10-bit mantissa)."""
    if mode == "tf32-native":
        monkeypatch.setenv("EINCONV_CTMA_TF32", "1")
        mode = "tf32"
    elif mode == "tf32-half":
        monkeypatch.setenv("EINCONV_CTMA_HALF", "1")
        mode = "tf32"
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "tf32": torch.float32}[mode]
    tol = {"bf16": 2.0**-7, "fp16": 2.0**-10, "tf32": 1e-3}[mode]
    x, w, b, kw = _tc_inputs(case, dtype)
    groups = kw.get("groups", 1)
    N = x.dim() - 2
    name = evaluator.selected_kernel  # noqa: F841 (kept for debugging sessions)
    want = direct.conv_forward(x.double().numpy(), w.double().numpy(), b.double().numpy(), **kw)
    torch.manual_seed(1)
    v = (torch.rand(*want.shape) - 0.5).to(dtype)
    want_gx = direct.conv_input_vjp(
        w.double().numpy(), v.double().numpy(), tuple(x.shape[2:]), **kw
    )
    evaluator.set_fp32_math("tf32" if mode == "tf32" else "auto")
    try:
        got = convNd(x.to(DEV), w.to(DEV), b.to(DEV), **kw)
        gx = evaluator.conv_input_vjp_raw(
            w.to(DEV), v.to(DEV), tuple(x.shape[2:]), kw.get("stride", 1), kw.get("padding", 0),
            kw.get("dilation", 1), groups,
        )
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, tol, f"conv_tma fwd {mode}")
    _rel_to_max(gx, want_gx, tol, f"conv_tma input vjp {mode}")


@pytest.mark.parametrize("case", CTMA_CASES, ids=CTMA_IDS)
@pytest.mark.parametrize("mode", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("taps_per_mma", ["auto", "pairs", "fused", "selfpack", "nb2"])
def test_slab_weight_vjp(case, mode, taps_per_mma, monkeypatch):
    """csrc/wgrad_slab.cu (MN-major operands from the channels-last padded copies, tap pairs sharing one
    slab): max|d| <= 2^-7 (bf16) / 2^-10 (fp16) of max|ref| against the fp64 oracle on the same bits;
    fp32 tensors in TF32 mode run on scaled fp16 operand copies with the TF32 tolerance 1e-3."""
    if taps_per_mma == "pairs":
        monkeypatch.setenv("EINCONV_WGRAD_NO_QUAD", "1")
    elif taps_per_mma == "fused":    # producer warps fill the slabs from the channels-first tensors (opt-in)
        monkeypatch.setenv("EINCONV_WGRAD_FUSED", "1")
    elif taps_per_mma == "selfpack":  # the kernel writes its own packed copies (forced on these small cases)
        monkeypatch.setenv("EINCONV_WGRAD_SELFPACK", "1")
    elif taps_per_mma == "nb2":      # two taps per unit (N = 128) where the dispatcher would take three or four
        monkeypatch.setenv("EINCONV_WGRAD_NB", "2")
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "tf32": torch.float32}[mode]
    x, w, b, kw = _tc_inputs(case, dtype)
    N = x.dim() - 2
    groups = kw.get("groups", 1)
    y = direct.conv_forward(x.double().numpy(), w.double().numpy(), None, **kw)
    torch.manual_seed(2)
    v = (torch.rand(*y.shape) - 0.5).to(dtype)
    want = direct.conv_weight_vjp(x.double().numpy(), v.double().numpy(), tuple(w.shape[2:]), **kw)
    evaluator.set_fp32_math("tf32" if mode == "tf32" else "auto")
    try:
        got = evaluator.conv_weight_vjp_raw(
            x.to(DEV), v.to(DEV), tuple(w.shape[2:]), kw.get("dilation", 1), kw.get("padding", 0),
            kw.get("stride", 1), groups,
        )
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, {"bf16": 2.0**-7, "fp16": 2.0**-10, "tf32": 1e-3}[mode], f"slab wgrad {mode}")


# ---- 1x1 convolutions: GEMMs on the channels-first tensors (csrc/conv_dense.cu), no packed copies ----------
DENSE_CASES = [
    ((2, 64, 16, 16), (64, 64, 1, 1)),       # one K block, N = 64
    ((3, 72, 8, 12), (200, 72, 1, 1)),       # K = 72 (partial second block), N = 200 of 208 / 256 columns
    ((2, 256, 8, 8), (520, 256, 1, 1)),      # four K blocks, three N tiles
    ((1, 136, 40), (24, 136, 1)),            # N = 1 (1-d), partial position tile
    ((2, 32, 4, 6, 8), (48, 32, 1, 1, 1)),   # N = 3: spatial dims collapse
    ((5, 8, 24, 24), (16, 8, 1, 1)),         # K = 8: mostly zero-filled K block; 4.5 position tiles per sample
]
DENSE_IDS = ["x".join(map(str, c[0])) + "-w" + "x".join(map(str, c[1])) for c in DENSE_CASES]


@pytest.mark.parametrize("case", DENSE_CASES, ids=DENSE_IDS)
@pytest.mark.parametrize("mode", ["bf16", "fp16", "tf32"])
def test_dense_1x1_forward_and_input_vjp(case, mode):
    """The dense rule's kernel: MN-major A operand fetched by TMA from the channels-first tensor, weights read
    in place (K-major for the forward pass, MN-major for the input VJP).  Same tolerances as the other
    tensor-core tests against the fp64 oracle on the same rounded inputs."""
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "tf32": torch.float32}[mode]
    tol = {"bf16": 2.0**-7, "fp16": 2.0**-10, "tf32": 1e-3}[mode]
    xs, ws = case
    torch.manual_seed(4)
    x = (torch.rand(*xs) - 0.3).to(dtype)
    w = (torch.rand(*ws) - 0.5).to(dtype)
    b = (torch.rand(ws[0]) - 0.5).to(dtype)
    N = len(xs) - 2
    want = direct.conv_forward(x.double().numpy(), w.double().numpy(), b.double().numpy(), stride=1, padding=0,
                               dilation=1, groups=1)
    v = (torch.rand(*want.shape) - 0.5).to(dtype)
    want_gx = direct.conv_input_vjp(w.double().numpy(), v.double().numpy(), tuple(xs[2:]), stride=1, padding=0,
                                    dilation=1, groups=1)
    g = evaluator._geom(xs[0], 1, xs[1], ws[0], tuple(xs[2:]), tuple(ws[2:]), 1, 0, 1)[0]
    evaluator.set_fp32_math("tf32" if mode == "tf32" else "auto")
    try:
        assert evaluator.selected_kernel(nat.OP_CONV_FWD, g, dtype) == "tcgen05_dense_conv_fwd"
        assert evaluator.selected_kernel(nat.OP_CONV_INPUT_VJP, g, dtype) == "tcgen05_dense_conv_input_vjp"
        got = convNd(x.to(DEV), w.to(DEV), b.to(DEV))
        gx = evaluator.conv_input_vjp_raw(w.to(DEV), v.to(DEV), tuple(xs[2:]), 1, 0, 1, 1)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, tol, f"dense 1x1 fwd {mode}")
    _rel_to_max(gx, want_gx, tol, f"dense 1x1 input vjp {mode}")


# ---- stride-1 2-d forward on fp32 tensors from the channels-first tensor (csrc/conv_rows.cu) -------------------
ROWS_CASES = [
    ((2, 64, 56, 56), (64, 64, 3, 3), dict(padding=1)),        # C1 geometry: two blocks per row, N = 192
    ((2, 40, 20, 24), (48, 40, 3, 3), dict(padding=1)),        # one block per row (4 rows per tile), K = 40
    ((1, 16, 9, 100), (80, 16, 3, 3), dict(padding=1)),        # four blocks per row, N = 240
    ((2, 8, 12, 72), (16, 8, 3, 3), dict(padding=(1, 1))),     # three blocks -> four (one empty)
    ((2, 24, 10, 32), (32, 24, 5, 4), dict(padding=(2, 1))),   # KW = 4 (N = 128), KH = 5, asymmetric result width
    ((3, 32, 7, 28), (64, 32, 1, 2), dict(padding=(0, 1))),    # KH = 1, KW = 2
    ((2, 16, 16, 40), (16, 16, 3, 3), dict(padding=0)),        # no padding: outputs narrower than the input
    ((2, 70, 6, 44), (30, 70, 2, 3), dict(padding=(1, 2))),    # padding = KW - 1, three K blocks (70 channels)
    ((3, 32, 12, 40), (200, 32, 3, 3), dict(padding=1)),       # three destination-channel classes (N = 3 * 80 each)
    ((2, 48, 9, 32), (100, 48, 3, 2), dict(padding=(1, 0))),   # KW = 2: one class of 112 would not fit, two of 64
]
ROWS_IDS = ["x".join(map(str, c[0])) + "-w" + "x".join(map(str, c[1])) for c in ROWS_CASES]


@pytest.mark.parametrize("variant", ["EINCONV_ROWS_PAIR=1", "EINCONV_ROWS_KSPLIT=2", "EINCONV_ROWS_PDL=1"])
@pytest.mark.parametrize("case", ROWS_CASES[:5], ids=ROWS_IDS[:5])
def test_rows_forward_variants(case, variant, monkeypatch):
    """Opt-in variants of the row kernel (CTA pairs with cta_group::2 MMAs, half-size pipeline stages, programmatic
    dependent launch) against the fp64 oracle, with bias."""
    xs, ws, kw = case
    torch.manual_seed(7)
    x = torch.rand(*xs) - 0.3
    w = torch.rand(*ws) - 0.5
    b = torch.rand(ws[0]) - 0.5
    want = direct.conv_forward(x.double().numpy(), w.double().numpy(), b.double().numpy(), stride=1, dilation=1,
                               groups=1, **kw)
    k, v = variant.split("=")
    monkeypatch.setenv(k, v)
    evaluator.set_fp32_math("tf32")
    try:
        got = convNd(x.to(DEV), w.to(DEV), b.to(DEV), **kw)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, 1e-3, f"rows fwd {variant}")


@pytest.mark.parametrize("case", ROWS_CASES, ids=ROWS_IDS)
@pytest.mark.parametrize("with_bias", [False, True])
def test_rows_forward_tf32(case, with_bias):
    """kh taps as TMA boxes of whole image rows, kw taps summed over neighbouring TMEM lanes in the epilogue:
    TF32 tolerance 1e-3 of max|ref| against the fp64 oracle."""
    xs, ws, kw = case
    torch.manual_seed(5)
    x = torch.rand(*xs) - 0.3
    w = torch.rand(*ws) - 0.5
    b = torch.rand(ws[0]) - 0.5 if with_bias else None
    want = direct.conv_forward(x.double().numpy(), w.double().numpy(), None if b is None else b.double().numpy(),
                               stride=1, dilation=1, groups=1, **kw)
    evaluator.set_fp32_math("tf32")
    try:
        g = evaluator._geom(xs[0], 1, xs[1], ws[0], tuple(xs[2:]), tuple(ws[2:]), 1, kw.get("padding", 0), 1)[0]
        assert evaluator.selected_kernel(nat.OP_CONV_FWD, g, torch.float32) == "tcgen05_rows_conv_fwd"
        got = convNd(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV), **kw)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, 1e-3, "rows fwd tf32")


@pytest.mark.parametrize("case", ROWS_CASES, ids=ROWS_IDS)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_rows_forward_and_input_vjp_16bit(case, dtype):
    """16-bit tensors on the row kernel: blocks of 32 positions in 64-byte rows (MN-major operand through the
    64-byte swizzle), so a block is still one warp's lanes.  Oracle in fp64 on the same rounded inputs; tolerance
    2^-7 (bf16) / 2^-10 (fp16) of max|ref| for the rounded output."""
    xs, ws, kw = case
    if (xs[3] * 2) % 16:
        pytest.skip("rows of the source are not 16-byte multiples in 16 bits")
    pad = kw.get("padding", 0)
    pad = (pad, pad) if isinstance(pad, int) else tuple(pad)
    out = tuple(xs[2 + d] + 2 * pad[d] - ws[2 + d] + 1 for d in range(2))
    tol = 2.0**-7 if dtype == torch.bfloat16 else 2.0**-10
    torch.manual_seed(8)
    x = (torch.rand(*xs) - 0.3).to(dtype)
    w = (torch.rand(*ws) - 0.5).to(dtype)
    b = (torch.rand(ws[0]) - 0.5).to(dtype)
    v = (torch.rand(xs[0], ws[0], *out) - 0.3).to(dtype)
    want = direct.conv_forward(x.double().numpy(), w.double().numpy(), b.double().numpy(), stride=1, dilation=1,
                               groups=1, padding=pad)
    want_gx = direct.conv_input_vjp(w.double().numpy(), v.double().numpy(), tuple(xs[2:]), stride=1, dilation=1,
                                    groups=1, padding=pad)
    g = evaluator._geom(xs[0], 1, xs[1], ws[0], tuple(xs[2:]), tuple(ws[2:]), 1, pad, 1)[0]
    assert evaluator.selected_kernel(nat.OP_CONV_FWD, g, dtype) == "tcgen05_rows_conv_fwd"
    got = convNd(x.to(DEV), w.to(DEV), b.to(DEV), padding=pad)
    _rel_to_max(got, want, tol, f"rows fwd {dtype}")
    if (out[1] * 2) % 16 == 0 and ws[0] <= 100:
        assert evaluator.selected_kernel(nat.OP_CONV_INPUT_VJP, g, dtype) == "tcgen05_rows_conv_input_vjp"
    gx = evaluator.conv_input_vjp(w.to(DEV), v.to(DEV), tuple(xs[2:]), padding=pad)
    _rel_to_max(gx, want_gx, tol, f"rows input vjp {dtype}")


@pytest.mark.parametrize("case", ROWS_CASES, ids=ROWS_IDS)
def test_rows_input_vjp_tf32(case):
    """The input VJP of the same geometries is the row kernel on the cotangent with flipped weights and
    padding k - 1 - p (weights flipped by the packing pre-pass): TF32 tolerance 1e-3 of max|ref|."""
    xs, ws, kw = case
    pad = kw.get("padding", 0)
    pad = (pad, pad) if isinstance(pad, int) else tuple(pad)
    out = tuple(xs[2 + d] + 2 * pad[d] - ws[2 + d] + 1 for d in range(2))
    torch.manual_seed(6)
    w = torch.rand(*ws) - 0.5
    v = torch.rand(xs[0], ws[0], *out) - 0.3
    want = direct.conv_input_vjp(w.double().numpy(), v.double().numpy(), tuple(xs[2:]), stride=1, dilation=1,
                                 groups=1, padding=pad)
    evaluator.set_fp32_math("tf32")
    try:
        g = evaluator._geom(xs[0], 1, xs[1], ws[0], tuple(xs[2:]), tuple(ws[2:]), 1, pad, 1)[0]
        # the cotangent is the TMA source here: its rows must be 16-byte multiples, and the flipped weights of all
        # C_out source channels must fit in shared memory (not for the 200-channel case)
        if (out[1] * 4) % 16 == 0 and ws[0] <= 100:
            assert evaluator.selected_kernel(nat.OP_CONV_INPUT_VJP, g, torch.float32) == "tcgen05_rows_conv_input_vjp"
        got = evaluator.conv_input_vjp(w.to(DEV), v.to(DEV), tuple(xs[2:]), padding=pad)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, 1e-3, "rows input vjp tf32")


def test_conv_tma_is_used_by_autograd_and_matches_torch():
    torch.manual_seed(0)
    x = torch.randn(4, 32, 18, 18, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    w = (torch.randn(48, 32, 3, 3, device=DEV) / 17).to(torch.bfloat16).requires_grad_(True)
    y = convNd(x, w, None, padding=1)
    ref = F.conv2d(x.float(), w.float(), None, padding=1)
    _rel_to_max(y, ref.detach().cpu().numpy(), 2.0**-7, "fwd vs torch")
    g = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, (x, w), g)
    rx, rw = torch.autograd.grad(ref, (x, w), g.float())
    _rel_to_max(gx, rx.detach().double().cpu().numpy(), 2.0**-7, "gx vs torch")
    _rel_to_max(gw, rw.detach().double().cpu().numpy(), 2.0**-6, "gw vs torch")


# ---- TMA-fed reduction kernels: geometries whose rows are 16-byte aligned ---------------------------
TMA_CASES = [
    ((2, 32, 24, 24), (48, 32, 3, 3), dict(padding=1)),
    ((2, 16, 16, 32), (24, 16, 3, 3), dict(padding=1, stride=2)),
    ((2, 24, 16, 16), (16, 12, 5, 3), dict(padding=(2, 0), dilation=(1, 2), groups=2)),
    ((2, 8, 256), (24, 8, 5), dict(padding=2)),
    ((1, 8, 6, 8, 16), (16, 8, 3, 3, 3), dict(padding=1)),
    ((2, 3, 32, 32), (16, 3, 7, 7), dict(padding=3, stride=2)),      # C_in = 3 (partial channel block), 49 taps
    ((2, 72, 8, 8), (200, 72, 1, 1), dict()),                         # BN = 256, c_in not a power of two
    ((1, 160, 8, 16), (32, 160, 3, 3), dict(padding=1)),              # two channel blocks of 128
    ((3, 64, 8, 8), (64, 64, 3, 3), dict(padding=1)),
]
TMA_IDS = ["x".join(map(str, c[0])) + "-w" + "x".join(map(str, c[1])) for c in TMA_CASES]


@pytest.mark.parametrize("case", TMA_CASES, ids=TMA_IDS)
@pytest.mark.parametrize("mode", ["bf16", "fp16", "tf32"])
def test_tma_weight_vjp(case, mode):
    xs, ws, kw = case
    torch.manual_seed(1)
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "tf32": torch.float32}[mode]
    x = (torch.rand(*xs) - 0.3).to(dtype)
    yshape = forward_out_shape(dict(x=xs, w=ws, kwargs=kw))
    v = (torch.rand(*yshape) - 0.5).to(dtype)
    want = direct.conv_weight_vjp(x.double().numpy(), v.double().numpy(), tuple(ws[2:]), **kw)
    evaluator.set_fp32_math("tf32" if mode == "tf32" else "auto")
    try:
        got = evaluator.conv_weight_vjp_raw(
            x.to(DEV), v.to(DEV), tuple(ws[2:]), kw.get("dilation", 1), kw.get("padding", 0),
            kw.get("stride", 1), kw.get("groups", 1),
        )
    finally:
        evaluator.set_fp32_math("auto")
    tol = {"bf16": 2.0**-7, "fp16": 2.0**-10, "tf32": 1e-3}[mode]
    _rel_to_max(got, want, tol, f"tma wgrad {mode}")


@pytest.mark.parametrize("case", TMA_CASES, ids=TMA_IDS)
@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_tma_kfc(case, mode):
    xs, ws, kw = case
    torch.manual_seed(2)
    dtype = torch.bfloat16 if mode == "bf16" else torch.float32
    x = torch.rand(*xs).to(dtype)
    want = direct.kfc(x.double().numpy(), tuple(ws[2:]), **kw)
    evaluator.set_fp32_math("tf32" if mode == "tf32" else "auto")
    try:
        got = evaluator.kfc(x.to(DEV), tuple(ws[2:]), **kw)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, 2.0**-7 if mode == "bf16" else 1e-3, f"tma kfc {mode}")


@pytest.mark.parametrize("case", TMA_CASES, ids=TMA_IDS)
@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_tensor_core_kfac_reduce(case, mode):
    """KFAC-reduce with the outer products on tensor cores (TF32 on the fp32 window sums):
    max|d| <= 1e-3 max|ref| (tf32) / 2^-7 (bf16 output rounding)."""
    xs, ws, kw = case
    torch.manual_seed(3)
    dtype = torch.bfloat16 if mode == "bf16" else torch.float32
    x = torch.rand(*xs).to(dtype)
    want = direct.kfac_reduce(x.double().numpy(), tuple(ws[2:]), **kw)
    evaluator.set_fp32_math("tf32" if mode == "tf32" else "auto")
    try:
        got = evaluator.kfac_reduce(x.to(DEV), tuple(ws[2:]), **kw)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, 2.0**-7 if mode == "bf16" else 1e-3, f"kfac_reduce {mode}")


# =============================================================================================
# strided forward on the TMA kernel (phase sub-images, csrc/conv_tma.cu Role::strided)
# =============================================================================================
CTMA_STRIDED = [
    ((2, 32, 20, 20), (48, 32, 3, 3), dict(padding=1, stride=2)),              # ResNet down-sampling 3x3
    ((2, 64, 14, 14), (128, 64, 1, 1), dict(stride=2)),                        # 1x1 stride 2: one phase only
    ((2, 3, 38, 37), (24, 3, 7, 7), dict(padding=3, stride=2)),                # stem: C_in = 3, 49 taps
    ((3, 16, 23, 17), (40, 16, 3, 3), dict(padding=1, stride=(2, 1))),         # mixed strides
    ((2, 24, 19, 21), (16, 12, 5, 3), dict(padding=(2, 0), dilation=(1, 2), stride=(3, 2), groups=2)),
    ((2, 8, 300), (24, 8, 7), dict(padding=3, stride=4)),                      # N = 1, stride > kernel/2
    ((1, 8, 9, 12, 10), (16, 8, 3, 3, 3), dict(padding=1, stride=2)),          # N = 3
    ((2, 16, 16, 16), (32, 16, 4, 4), dict(stride=4)),                         # patchify: K == S, one tap per phase
    ((2, 16, 15, 15), (32, 16, 2, 2), dict(stride=3, padding=1)),              # stride > kernel: skipped inputs
]


@pytest.mark.parametrize("case", CTMA_STRIDED, ids=["x".join(map(str, c[0])) + "-w" + "x".join(map(str, c[1])) for c in CTMA_STRIDED])
@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_conv_tma_strided_forward(case, mode):
    dtype = {"bf16": torch.bfloat16, "tf32": torch.float32}[mode]
    tol = {"bf16": 2.0**-7, "tf32": 1e-3}[mode]
    x, w, b, kw = _tc_inputs(case, dtype)
    want = direct.conv_forward(x.double().numpy(), w.double().numpy(), b.double().numpy(), **kw)
    evaluator.set_fp32_math("tf32" if mode == "tf32" else "auto")
    try:
        B, groups = x.shape[0], kw.get("groups", 1)
        plan = evaluator.call_plan(nat.OP_CONV_FWD, B, groups, w.shape[1], w.shape[0] // groups, x.shape[2:],
                                   tuple(w.shape[2:]), kw.get("stride", 1), kw.get("padding", 0),
                                   kw.get("dilation", 1), dtype)
        # (grouped layers whose per-group channel block is not 16-byte aligned keep the software-gather kernel)
        aligned = groups == 1 or (w.shape[1] * x.element_size()) % 16 == 0
        assert plan.kernel == ("tcgen05_tma_conv_fwd" if aligned else plan.kernel), plan.kernel
        got = convNd(x.to(DEV), w.to(DEV), b.to(DEV), **kw)
        # input VJP of the same layer: one launch per output phase over the packed cotangent
        torch.manual_seed(1)
        v = (torch.rand(*want.shape) - 0.5).to(dtype)
        want_gx = direct.conv_input_vjp(w.double().numpy(), v.double().numpy(), tuple(x.shape[2:]), **kw)
        vplan = evaluator.call_plan(nat.OP_CONV_INPUT_VJP, B, groups, w.shape[1], w.shape[0] // groups, x.shape[2:],
                                    tuple(w.shape[2:]), kw.get("stride", 1), kw.get("padding", 0),
                                    kw.get("dilation", 1), dtype)
        aligned_v = groups == 1 or ((w.shape[0] // groups) * x.element_size()) % 16 == 0
        assert vplan.kernel == ("tcgen05_tma_conv_input_vjp" if aligned_v else vplan.kernel), vplan.kernel
        gx = evaluator.conv_input_vjp_raw(w.to(DEV), v.to(DEV), tuple(x.shape[2:]), kw.get("stride", 1),
                                          kw.get("padding", 0), kw.get("dilation", 1), groups)
    finally:
        evaluator.set_fp32_math("auto")
    _rel_to_max(got, want, tol, f"strided conv_tma fwd {mode}")
    _rel_to_max(gx, want_gx, tol, f"strided conv_tma input vjp {mode}")
