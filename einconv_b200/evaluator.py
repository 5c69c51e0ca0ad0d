"""Native evaluation of einconv's convolution tensor networks on sm_100a.

This module is the host side of the drop-in boundary: it stands where the reference calls
``torch.einsum(equation, *operands).reshape(shape)`` (``einconv/functionals/conv.py:61``,
``einconv/functionals/unfold.py:56``, user code per ``README.md:64-73``) and calls the C ABI of
``libeinconv_b200.so`` instead.  Geometry is resolved exactly as the reference does
(``einconv/utils.py:97-160``) and handed to the kernels as an ``einconv_geom``; index patterns are
never built.

Two entry styles:

* direct functions (``conv_forward``, ``conv_input_vjp``, ``conv_weight_vjp``, ``unfold``, ``kfc``,
  ``kfac_reduce``) used by ``functionals`` / ``modules``;
* :func:`einsum`, reached through ``torch.einsum`` whenever an operand is a lazy
  :class:`~einconv_b200.conv_index_pattern.IndexPattern`: expressions returned by
  ``einconv_b200.expressions.*.einsum_expression`` carry a :class:`Plan` on their equation string,
  so the familiar ``torch.einsum(equation, *operands).reshape(shape)`` runs the native kernel.
"""

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from einconv_b200 import _native as nat
from einconv_b200.utils import _tuple, get_conv_input_size, resolve_geometry

# --------------------------------------------------------------------------------------------
# math-mode policy
# --------------------------------------------------------------------------------------------
_FP32_MATH = "auto"


def set_fp32_math(mode: str) -> None:
    """How float32 contractions are computed.

    ``"fp32"``: CUDA-core FMA, full fp32 accumulation (meets the reference's ``rtol=1e-5``).
    ``"tf32"``: tcgen05 ``kind::tf32`` tensor cores with fp32 accumulation in TMEM.
    ``"auto"`` (default): follow ``torch.backends.cuda.matmul.allow_tf32`` — the switch that
    governs the ``bmm`` calls the reference's ``torch.einsum`` ends in.
    """
    global _FP32_MATH
    if mode not in ("auto", "fp32", "tf32"):
        raise ValueError(f"Unknown fp32 math mode {mode!r}.")
    _FP32_MATH = mode


def get_fp32_math() -> str:
    return _FP32_MATH


def _math_for(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        mode = _FP32_MATH
        if mode == "auto":
            mode = "tf32" if torch.backends.cuda.matmul.allow_tf32 else "fp32"
        return nat.MATH_TF32 if mode == "tf32" else nat.MATH_FP32
    return nat.MATH_AUTO


def _kernels_for(simplify: bool) -> int:
    # simplify=True may use specialised kernels, simplify=False always the generic one; results
    # agree (the reference tests every case both ways, test/utils.py:98).
    return nat.KERNELS_SPECIALIZED if simplify else nat.KERNELS_GENERIC


# --------------------------------------------------------------------------------------------
# geometry helpers
# --------------------------------------------------------------------------------------------
def _check_sizes(out: Sequence[int], what: str) -> None:
    # The reference fails inside conv1d while building the pattern (conv_index_pattern.py:61):
    # "Kernel size can't be greater than actual input size".  Same exception type here.
    if any(o <= 0 for o in out):
        raise RuntimeError(
            f"{what}: kernel size can't be greater than the padded input size "
            f"(output sizes would be {tuple(out)})."
        )


def _hashable(h):
    return tuple(h) if isinstance(h, list) else h


class CallPlan:
    """Everything about one (operation, geometry, dtype, math mode, kernel family) that does not depend on
    the tensors' contents: the resolved ``einconv_geom``, output sizes, workspace bytes and (lazily) the
    name of the kernel the dispatcher picks.

    This is the tensor-free plan the reference's symbolic simplifier aims at
    (``einconv/simplifications/symbolic.py:518-544``, ``README.md:110-117``): the reference re-derives
    equation, patterns and simplifications from the tensors at every call; here a repeated call costs one
    dictionary lookup before the kernel launch."""

    __slots__ = ("op", "geom", "kernel_size", "out", "code", "math", "kern", "ws_bytes", "_kernel")

    def __init__(self, op, geom, kernel_size, out, code, math, kern):
        self.op, self.geom, self.kernel_size, self.out = op, geom, kernel_size, out
        self.code, self.math, self.kern = code, math, kern
        self.ws_bytes = 0
        self._kernel = None

    @property
    def kernel(self) -> str:
        if self._kernel is None:
            name = nat.lib().einconv_selected_kernel(self.op, self.geom, self.code, self.math, self.kern)
            self._kernel = name.decode() if name else ""
        return self._kernel

    def workspace(self, device) -> Tuple[Optional[Tensor], int]:
        if self.ws_bytes == 0:
            return None, 0
        return torch.empty(self.ws_bytes, dtype=torch.uint8, device=device), self.ws_bytes


_PLANS: Dict[tuple, CallPlan] = {}
_PLAN_STATS = {"hits": 0, "misses": 0}
_PLAN_CAPACITY = 8192


def plan_cache_info() -> Dict[str, int]:
    """``{"hits", "misses", "size"}`` of the call-plan cache."""
    return dict(_PLAN_STATS, size=len(_PLANS))


def clear_plan_cache() -> None:
    _PLANS.clear()
    _PLAN_STATS["hits"] = _PLAN_STATS["misses"] = 0


def call_plan(op, batch, groups, c_in_g, c_out_g, input_size, kernel_size, stride, padding, dilation, dtype,
              simplify=True) -> CallPlan:
    """Cached :class:`CallPlan`; geometry errors are raised (not cached) exactly as without the cache."""
    math = _math_for(dtype)
    key = (op, batch, groups, c_in_g, c_out_g, tuple(input_size), _hashable(kernel_size), _hashable(stride),
           _hashable(padding), _hashable(dilation), dtype, math, simplify)
    plan = _PLANS.get(key)
    if plan is not None:
        _PLAN_STATS["hits"] += 1
        return plan
    _PLAN_STATS["misses"] += 1
    t_in, t_k, t_s, t_d, p_left, out = resolve_geometry(input_size, kernel_size, stride, padding, dilation)
    _check_sizes(out, "convolution geometry")
    g = nat.make_geom(batch, groups, c_in_g, c_out_g, t_in, t_k, t_s, t_d, p_left, out)
    plan = CallPlan(op, g, t_k, out, nat.dtype_code(dtype), math, _kernels_for(simplify))
    if op not in (nat.OP_UNFOLD, nat.OP_FOLD, nat.OP_UNFOLD_TRANSPOSE):
        nbytes = nat.lib().einconv_workspace_bytes(op, g, plan.code, plan.math, plan.kern)
        if nbytes < 0:
            nat.check(-1, "einconv_workspace_bytes")
        plan.ws_bytes = int(nbytes)
    if len(_PLANS) >= _PLAN_CAPACITY:
        _PLANS.pop(next(iter(_PLANS)))
    _PLANS[key] = plan
    return plan


def _geom(batch, groups, c_in_g, c_out_g, input_size, kernel_size, stride, padding, dilation):
    """Uncached geometry (tests and tools): ``(einconv_geom, kernel sizes, output sizes)``."""
    t_in, t_k, t_s, t_d, p_left, out = resolve_geometry(
        input_size, kernel_size, stride, padding, dilation
    )
    _check_sizes(out, "convolution geometry")
    g = nat.make_geom(batch, groups, c_in_g, c_out_g, t_in, t_k, t_s, t_d, p_left, out)
    return g, t_k, out


class _on_device:
    """``torch.cuda.device(device)`` only when it is not already current (saves ~3 us per call)."""

    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index is None or torch.cuda.current_device() == device.index \
            else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def _workspace(op: int, g, dtype_code: int, math: int, kernels: int, device) -> Tuple[Optional[Tensor], int]:
    nbytes = nat.lib().einconv_workspace_bytes(op, g, dtype_code, math, kernels)
    if nbytes < 0:
        nat.check(-1, "einconv_workspace_bytes")
    if nbytes == 0:
        return None, 0
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def _same_dtype(*tensors: Tensor) -> torch.dtype:
    dtype = tensors[0].dtype
    for t in tensors[1:]:
        if t is not None and t.dtype != dtype:
            raise RuntimeError(f"Expected operands of equal dtype. Got {dtype} and {t.dtype}.")
    return dtype


def selected_kernel(op: int, g, dtype: torch.dtype, simplify: bool = True) -> str:
    """Name of the kernel family the dispatcher picks (for tests and logs)."""
    name = nat.lib().einconv_selected_kernel(
        op, g, nat.dtype_code(dtype), _math_for(dtype), _kernels_for(simplify)
    )
    return name.decode() if name else ""


# --------------------------------------------------------------------------------------------
# raw (non-differentiable) native calls
# --------------------------------------------------------------------------------------------
def unfold(x: Tensor, kernel_size, dilation=1, padding=0, stride=1) -> Tensor:
    """``[B, C, *in] -> [B, C * prod(k), prod(out)]`` — network of ``convNd_unfold.py:42-53``."""
    device = nat.require_cuda(x)
    if x.dim() < 3:
        raise ValueError(f"unfoldNd expects [batch, channels, *spatial]. Got shape {tuple(x.shape)}.")
    B, C = x.shape[:2]
    plan = call_plan(nat.OP_UNFOLD, B, 1, C, 0, x.shape[2:], kernel_size, stride, padding, dilation, x.dtype)
    g, t_k, out = plan.geom, plan.kernel_size, plan.out
    k_total = 1
    for k in t_k:
        k_total *= k
    o_total = 1
    for o in out:
        o_total *= o
    result = torch.empty((B, C * k_total, o_total), dtype=x.dtype, device=device)
    if result.numel() == 0:
        return result
    strides = None if x.is_contiguous() else (nat.ctypes.c_int64 * x.dim())(*x.stride())
    with _on_device(device):
        st = nat.lib().einconv_unfold(
            x.data_ptr(), strides, result.data_ptr(), plan.code, g, nat.stream_ptr(device)
        )
    nat.check(st, "einconv_unfold")
    return result


def fold(gu: Tensor, channels: int, input_size, kernel_size, dilation=1, padding=0, stride=1) -> Tensor:
    """Adjoint of :func:`unfold`: ``[B, C * prod(k), prod(out)] -> [B, C, *in]``."""
    device = nat.require_cuda(gu)
    gu = gu.contiguous()
    B = gu.shape[0]
    g, _, _ = _geom(B, 1, channels, 0, input_size, kernel_size, stride, padding, dilation)
    result = torch.empty((B, channels, *input_size), dtype=gu.dtype, device=device)
    if result.numel() == 0:
        return result
    with torch.cuda.device(device):
        st = nat.lib().einconv_fold(
            gu.data_ptr(), result.data_ptr(), nat.dtype_code(gu.dtype), g, nat.stream_ptr(device)
        )
    nat.check(st, "einconv_fold")
    return result


def _conv_check(x: Tensor, weight: Tensor, bias: Optional[Tensor], groups: int) -> None:
    """Argument checks of ``einconv/functionals/conv.py:72-116`` (same errors)."""
    N = x.dim() - 2
    in_channels = x.shape[1]
    out_channels = weight.shape[0]
    if weight.dim() != N + 2:
        raise ValueError(f"For Conv(N={N})d, the kernel must be {N+2}d. Got {weight.dim()}d.")
    if weight.shape[0] % groups != 0:
        raise ValueError(f"Groups ({groups}) must divide out_channels ({weight.shape[0]}).")
    if weight.shape[1] * groups != in_channels:
        raise ValueError(
            f"Kernel dimension 1 ({weight.shape[1]}) multiplied by groups ({groups})"
            + f" must equal in_channels ({in_channels})."
        )
    if bias is not None and (bias.dim() != 1 or bias.numel() != out_channels):
        raise ValueError(f"Bias should have shape [{out_channels}]. Got {bias.shape}.")


def conv_forward_raw(x, weight, bias, stride, padding, dilation, groups, simplify=True) -> Tensor:
    device = nat.require_cuda(x, weight, bias)
    dtype = _same_dtype(x, weight, bias)
    x, weight = x.contiguous(), weight.contiguous()
    bias = None if bias is None else bias.contiguous()
    B, c_out = x.shape[0], weight.shape[0]
    plan = call_plan(nat.OP_CONV_FWD, B, groups, weight.shape[1], c_out // groups, x.shape[2:],
                     tuple(weight.shape[2:]), stride, padding, dilation, dtype, simplify)
    g, out = plan.geom, plan.out
    y = torch.empty((B, c_out, *out), dtype=dtype, device=device)
    if y.numel() == 0:
        return y
    code, math, kern = plan.code, plan.math, plan.kern
    ws, ws_bytes = plan.workspace(device)
    with _on_device(device):
        st = nat.lib().einconv_conv_fwd(
            x.data_ptr(), weight.data_ptr(), nat.ptr(bias), y.data_ptr(), code, g, math, kern,
            nat.ptr(ws), ws_bytes, nat.stream_ptr(device),
        )
    nat.check(st, "einconv_conv_fwd")
    return y


def conv_input_vjp_raw(weight, v, input_size, stride, padding, dilation, groups, simplify=True) -> Tensor:
    device = nat.require_cuda(weight, v)
    dtype = _same_dtype(weight, v)
    weight, v = weight.contiguous(), v.contiguous()
    N = weight.dim() - 2
    t_input = _tuple(input_size if isinstance(input_size, int) else tuple(input_size), N)
    B, c_out = v.shape[0], weight.shape[0]
    c_in_g = weight.shape[1]
    plan = call_plan(nat.OP_CONV_INPUT_VJP, B, groups, c_in_g, c_out // groups, t_input, tuple(weight.shape[2:]),
                     stride, padding, dilation, dtype, simplify)
    g, out = plan.geom, plan.out
    if tuple(v.shape[2:]) != tuple(out) or v.shape[1] != c_out:
        raise ValueError(
            f"Cotangent has shape {tuple(v.shape)}, expected {(B, c_out, *out)} for this geometry."
        )
    gx = torch.empty((B, groups * c_in_g, *t_input), dtype=dtype, device=device)
    if gx.numel() == 0:
        return gx
    if v.numel() == 0:
        return gx.zero_()
    code, math, kern = plan.code, plan.math, plan.kern
    ws, ws_bytes = plan.workspace(device)
    with _on_device(device):
        st = nat.lib().einconv_conv_input_vjp(
            weight.data_ptr(), v.data_ptr(), gx.data_ptr(), code, g, math, kern,
            nat.ptr(ws), ws_bytes, nat.stream_ptr(device),
        )
    nat.check(st, "einconv_conv_input_vjp")
    return gx


def conv_weight_vjp_raw(x, v, kernel_size, dilation, padding, stride, groups, simplify=True) -> Tensor:
    device = nat.require_cuda(x, v)
    dtype = _same_dtype(x, v)
    x, v = x.contiguous(), v.contiguous()
    N = x.dim() - 2
    B, c_in, c_out = x.shape[0], x.shape[1], v.shape[1]
    plan = call_plan(nat.OP_CONV_WEIGHT_VJP, B, groups, c_in // groups, c_out // groups, x.shape[2:], kernel_size,
                     stride, padding, dilation, dtype, simplify)
    g, t_k, out = plan.geom, plan.kernel_size, plan.out
    if tuple(v.shape[2:]) != tuple(out) or v.shape[0] != B:
        raise ValueError(
            f"Cotangent has shape {tuple(v.shape)}, expected {(B, c_out, *out)} for this geometry."
        )
    gw = torch.empty((c_out, c_in // groups, *t_k), dtype=dtype, device=device)
    if gw.numel() == 0:
        return gw
    code, math, kern = plan.code, plan.math, plan.kern
    ws, ws_bytes = plan.workspace(device)
    with _on_device(device):
        st = nat.lib().einconv_conv_weight_vjp(
            x.data_ptr(), v.data_ptr(), gw.data_ptr(), code, g, math, kern,
            nat.ptr(ws), ws_bytes, nat.stream_ptr(device),
        )
    nat.check(st, "einconv_conv_weight_vjp")
    return gw


def _factor(op: int, x, kernel_size, stride, padding, dilation, groups, simplify, scale, dest=None) -> Tensor:
    device = nat.require_cuda(x)
    x = x.contiguous()
    B, c_in = x.shape[0], x.shape[1]
    if c_in % groups != 0:
        raise ValueError(f"Groups ({groups}) must divide in_channels ({c_in}).")
    plan = call_plan(op, B, groups, c_in // groups, 0, x.shape[2:], kernel_size, stride, padding, dilation, x.dtype,
                     simplify)
    g, t_k, out = plan.geom, plan.kernel_size, plan.out
    k_total = 1
    for k in t_k:
        k_total *= k
    A = c_in // groups * k_total
    if scale is None:
        if op == nat.OP_KFC:
            # Tensor([1 / B]) in x.dtype, convNd_kfc.py:105
            scale = float(torch.tensor([1.0 / B]).to(x.dtype)) if B > 0 else 0.0
        else:
            o_total = 1
            for o in out:
                o_total *= o
            # Tensor([1 / (B * prod(out)^2)]) in x.dtype, convNd_kfac_reduce.py:106-108
            denom = B * o_total**2
            scale = float(torch.tensor([1.0 / denom]).to(x.dtype)) if denom > 0 else 0.0
    if dest is None:
        result = torch.empty((groups, A, A), dtype=x.dtype, device=device)
    else:
        # caller-owned output (e.g. a slice of a flat all-reduce bucket, einconv_b200/distributed.py)
        if tuple(dest.shape) != (groups, A, A) or dest.dtype != x.dtype or not dest.is_contiguous() \
                or dest.device != x.device:
            raise ValueError(f"out must be a contiguous {x.dtype} tensor of shape {(groups, A, A)} on {x.device}.")
        result = dest
    if result.numel() == 0:
        return result
    code, math, kern = plan.code, plan.math, plan.kern
    ws, ws_bytes = plan.workspace(device)
    fn = nat.lib().einconv_kfc if op == nat.OP_KFC else nat.lib().einconv_kfac_reduce
    with _on_device(device):
        st = fn(
            x.data_ptr(), result.data_ptr(), code, g, float(scale), math, kern,
            nat.ptr(ws), ws_bytes, nat.stream_ptr(device),
        )
    nat.check(st, "einconv_kfc" if op == nat.OP_KFC else "einconv_kfac_reduce")
    return result


def kfc(x, kernel_size, stride=1, padding=0, dilation=1, groups=1, simplify=True, scale=None, out=None) -> Tensor:
    """KFC factor ``[G, A, A]`` (``convNd_kfc.py:75-116``).  ``scale`` defaults to ``1/B``; pass the
    global ``1/B`` when ``x`` is a batch shard whose factors will be summed across ranks.  ``out``
    (optional) receives the result in place."""
    return _factor_entry(nat.OP_KFC, x, kernel_size, stride, padding, dilation, groups, simplify, scale, out)


def kfac_reduce(x, kernel_size, stride=1, padding=0, dilation=1, groups=1, simplify=True, scale=None,
                out=None) -> Tensor:
    """KFAC-reduce factor ``[G, A, A]`` (``convNd_kfac_reduce.py:77-119``)."""
    return _factor_entry(nat.OP_KFAC_REDUCE, x, kernel_size, stride, padding, dilation, groups, simplify, scale, out)


def _factor_entry(op, x, kernel_size, stride, padding, dilation, groups, simplify, scale, out):
    if _needs_grad(x):
        if out is not None:
            raise RuntimeError("kfc / kfac_reduce: `out=` cannot be combined with an input that requires grad.")
        return _FactorFunction.apply(x, op, kernel_size, stride, padding, dilation, groups, simplify, scale)
    return _factor(op, x, kernel_size, stride, padding, dilation, groups, simplify, scale, out)


# --------------------------------------------------------------------------------------------
# KFAC-reduce in two stages (batch-sharded evaluation): window sums, then their Gram matrix
# --------------------------------------------------------------------------------------------
_STAGE2_WS: Dict[tuple, int] = {}


def kfac_reduce_sums(x, kernel_size, stride=1, padding=0, dilation=1, groups=1, out=None) -> Tensor:
    """Stage 1 of KFAC-reduce: ``u[g * A + a, b] = sum_o U[b, g, a, o]`` as ``[G * A, B]`` (fp32; fp64 for
    fp64 input), the unfolded input summed over output positions without ever being written
    (``convNd_kfac_reduce.py:77-108``: the network factorises over these sums)."""
    device = nat.require_cuda(x)
    x = x.contiguous()
    B, c_in = x.shape[0], x.shape[1]
    if c_in % groups != 0:
        raise ValueError(f"Groups ({groups}) must divide in_channels ({c_in}).")
    plan = call_plan(nat.OP_KFAC_REDUCE, B, groups, c_in // groups, 0, x.shape[2:], kernel_size, stride, padding,
                     dilation, x.dtype, True)
    k_total = 1
    for k in plan.kernel_size:
        k_total *= k
    rows = c_in * k_total
    u_dtype = torch.float64 if x.dtype == torch.float64 else torch.float32
    if out is None:
        out = torch.empty((rows, B), dtype=u_dtype, device=device)
    elif tuple(out.shape) != (rows, B) or out.dtype != u_dtype or not out.is_contiguous() or out.device != x.device:
        raise ValueError(f"out must be a contiguous {u_dtype} tensor of shape {(rows, B)} on {x.device}.")
    if out.numel() == 0:
        return out
    with _on_device(device):
        st = nat.lib().einconv_kfac_reduce_sums(x.data_ptr(), out.data_ptr(), plan.code, plan.geom,
                                                nat.stream_ptr(device))
    nat.check(st, "einconv_kfac_reduce_sums")
    return out


def kfac_reduce_from_sums_supported(plan: CallPlan, n_shards: int) -> bool:
    """Whether stage 2 can run on ``n_shards`` gathered blocks of window sums for this call plan."""
    return _stage2_workspace(plan, n_shards) >= 0


def _stage2_workspace(plan: CallPlan, n_shards: int) -> int:
    key = (n_shards, plan.code, plan.math, plan.kern, bytes(plan.geom))
    ws = _STAGE2_WS.get(key)
    if ws is None:
        ws = int(nat.lib().einconv_kfac_reduce_from_sums_workspace(n_shards, plan.code, plan.geom, plan.math,
                                                                     plan.kern))
        if len(_STAGE2_WS) > _PLAN_CAPACITY:
            _STAGE2_WS.clear()
        _STAGE2_WS[key] = ws
    return ws


def kfac_reduce_from_sums(u: Tensor, plan: CallPlan, n_shards: int, shard_stride: int, scale: float,
                          out: Tensor) -> Tensor:
    """Stage 2: ``out[g, a, a'] = scale * sum_s sum_b u_s[g, a, b] u_s[g, a', b]`` over ``n_shards`` blocks of
    window sums, block ``s`` starting ``s * shard_stride`` elements into ``u`` (e.g. the result of an
    all-gather over ranks), each ``[G * A, B_shard]`` with ``B_shard = plan.geom.batch``.  ``plan`` is the
    KFAC-reduce call plan of ONE shard (``call_plan(OP_KFAC_REDUCE, B_shard, ...)``)."""
    device = nat.require_cuda(u, out)
    ws_bytes = _stage2_workspace(plan, n_shards)
    if ws_bytes < 0:
        raise NotImplementedError("kfac_reduce_from_sums: no tensor-core route for this geometry / math mode")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device) if ws_bytes else None
    with _on_device(device):
        st = nat.lib().einconv_kfac_reduce_from_sums(
            u.data_ptr(), n_shards, shard_stride, out.data_ptr(), plan.code, plan.geom, float(scale), plan.math,
            plan.kern, nat.ptr(ws), ws_bytes, nat.stream_ptr(device))
    nat.check(st, "einconv_kfac_reduce_from_sums")
    return out


# --------------------------------------------------------------------------------------------
# transpose convolutions.  Hyper-parameters are those of the ASSOCIATED
# convolution; ``x`` has that convolution's OUTPUT sizes.
# --------------------------------------------------------------------------------------------
def _transpose_geom(x, groups, kernel_size, stride, padding, output_padding, dilation):
    """Geometry of the associated convolution for a transpose-convolution input ``x``
    (``conv_transposeNd_unfold.py:93-112``): its input sizes follow from ``x``'s spatial sizes."""
    N = x.dim() - 2
    t_k, t_s, t_p = _tuple(kernel_size, N), _tuple(stride, N), _tuple(padding, N)
    t_d, t_op = _tuple(dilation, N), _tuple(output_padding, N)
    in_sizes = tuple(
        get_conv_input_size(o, k, s, p, op, d)
        for o, k, s, p, op, d in zip(x.shape[2:], t_k, t_s, t_p, t_op, t_d)
    )
    c = x.shape[1]
    if c % groups != 0:
        raise ValueError(f"Groups ({groups}) must divide the number of channels ({c}).")
    g, _, out = _geom(x.shape[0], groups, c // groups, 0, in_sizes, t_k, t_s, t_p, t_d)
    assert tuple(out) == tuple(x.shape[2:])
    return g, t_k, in_sizes


def unfold_transpose(x: Tensor, kernel_size, stride=1, padding=0, output_padding=0, dilation=1) -> Tensor:
    """``[B, C, *out] -> [B, C * prod(k), prod(in)]`` — network of ``conv_transposeNd_unfold.py:81-91``
    (bit-exact gather ``out[b, c, k, i] = x[b, c, o]`` where ``i = s o + d k - p``)."""
    device = nat.require_cuda(x)
    x = x.contiguous()
    g, t_k, in_sizes = _transpose_geom(x, 1, kernel_size, stride, padding, output_padding, dilation)
    k_total = i_total = 1
    for k in t_k:
        k_total *= k
    for i in in_sizes:
        i_total *= i
    result = torch.empty((x.shape[0], x.shape[1] * k_total, i_total), dtype=x.dtype, device=device)
    if result.numel() == 0:
        return result
    with torch.cuda.device(device):
        st = nat.lib().einconv_unfold_transpose(
            x.data_ptr(), result.data_ptr(), nat.dtype_code(x.dtype), g, nat.stream_ptr(device)
        )
    nat.check(st, "einconv_unfold_transpose")
    return result


def kfac_reduce_transpose(x, kernel_size, stride=1, padding=0, output_padding=0, dilation=1, groups=1,
                          simplify=True, scale=None) -> Tensor:
    """Input-based KFAC-reduce factor ``[G, A, A]`` of a transpose convolution
    (``conv_transposeNd_kfac_reduce.py:95-160``); ``scale`` defaults to ``1 / (B * prod(in)^2)``."""
    device = nat.require_cuda(x)
    x = x.contiguous()
    g, t_k, in_sizes = _transpose_geom(x, groups, kernel_size, stride, padding, output_padding, dilation)
    k_total = i_total = 1
    for k in t_k:
        k_total *= k
    for i in in_sizes:
        i_total *= i
    B = x.shape[0]
    A = x.shape[1] // groups * k_total
    if _needs_grad(x):
        # differentiable by composition: native (differentiable) gather, then the definition
        if scale is None:
            denom = B * i_total**2
            scale = float(torch.tensor([1.0 / denom]).to(x.dtype)) if denom > 0 else 0.0
        u = unfold_transpose_autograd(x, kernel_size, stride, padding, output_padding, dilation)
        m = u.reshape(B, groups, A, i_total).sum(-1)
        return torch.einsum("ngi,ngj->gij", m, m) * scale
    if scale is None:
        denom = B * i_total**2
        scale = float(torch.tensor([1.0 / denom]).to(x.dtype)) if denom > 0 else 0.0
    result = torch.empty((groups, A, A), dtype=x.dtype, device=device)
    if result.numel() == 0:
        return result
    code, math, kern = nat.dtype_code(x.dtype), _math_for(x.dtype), _kernels_for(simplify)
    ws, ws_bytes = _workspace(nat.OP_KFAC_REDUCE_TRANSPOSE, g, code, math, kern, device)
    with torch.cuda.device(device):
        st = nat.lib().einconv_kfac_reduce_transpose(
            x.data_ptr(), result.data_ptr(), code, g, float(scale), math, kern,
            nat.ptr(ws), ws_bytes, nat.stream_ptr(device),
        )
    nat.check(st, "einconv_kfac_reduce_transpose")
    return result


# --------------------------------------------------------------------------------------------
# differentiable wrappers.  The reference evaluates every network with ``torch.einsum``, which is
# differentiable to any order (``modules/conv.py:173-182``, ``README.md:64-73``).  The kernels are
# each other's derivatives, so every backward below is written with the differentiable wrappers
# again: with ``create_graph=True`` the graph of the backward pass is recorded like einsum's would be.
#   y  = fwd(x, w):        dx = ivjp(w, c)        dw = wvjp(x, c)
#   gx = ivjp(w, v):       dw = wvjp(c, v)        dv = fwd(c, w)
#   gw = wvjp(x, v):       dx = ivjp(c, v)        dv = fwd(x, c)
#   u  = unfold(x):        dx = fold(c)           (and fold's backward is unfold)
#   K  = s sum_n U U^T:    dU = s (C + C^T) U     dx = fold(dU)
#   R  = s sum_n m m^T:    dm = s (C + C^T) m,  m = sum_o U,  dU = dm broadcast over o,  dx = fold(dU)
# --------------------------------------------------------------------------------------------
def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class _ConvNdFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups, simplify):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.hyper = (stride, padding, dilation, groups, simplify)
        return conv_forward_raw(x, weight, bias, stride, padding, dilation, groups, simplify)

    @staticmethod
    def backward(ctx, v):
        x, weight = ctx.saved_tensors
        stride, padding, dilation, groups, simplify = ctx.hyper
        gx = gw = gb = None
        v = v.contiguous()
        if ctx.needs_input_grad[0]:
            gx = conv_input_vjp(weight, v, tuple(x.shape[2:]), stride, padding, dilation, groups, simplify)
        if ctx.needs_input_grad[1]:
            gw = conv_weight_vjp(x, v, tuple(weight.shape[2:]), dilation, padding, stride, groups, simplify)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = v.sum(dim=[0] + list(range(2, v.dim())))
        return gx, gw, gb, None, None, None, None, None


class _InputVjpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, v, input_size, stride, padding, dilation, groups, simplify):
        ctx.save_for_backward(weight, v)
        ctx.hyper = (stride, padding, dilation, groups, simplify)
        return conv_input_vjp_raw(weight, v, input_size, stride, padding, dilation, groups, simplify)

    @staticmethod
    def backward(ctx, c):
        weight, v = ctx.saved_tensors
        stride, padding, dilation, groups, simplify = ctx.hyper
        gw = gv = None
        c = c.contiguous()
        if ctx.needs_input_grad[0]:
            gw = conv_weight_vjp(c, v, tuple(weight.shape[2:]), dilation, padding, stride, groups, simplify)
        if ctx.needs_input_grad[1]:
            gv = conv_forward(c, weight, None, stride, padding, dilation, groups, simplify)
        return gw, gv, None, None, None, None, None, None


class _WeightVjpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, v, kernel_size, dilation, padding, stride, groups, simplify):
        ctx.save_for_backward(x, v)
        ctx.hyper = (stride, padding, dilation, groups, simplify)
        return conv_weight_vjp_raw(x, v, kernel_size, dilation, padding, stride, groups, simplify)

    @staticmethod
    def backward(ctx, c):
        x, v = ctx.saved_tensors
        stride, padding, dilation, groups, simplify = ctx.hyper
        gx = gv = None
        c = c.contiguous()
        if ctx.needs_input_grad[0]:
            gx = conv_input_vjp(c, v, tuple(x.shape[2:]), stride, padding, dilation, groups, simplify)
        if ctx.needs_input_grad[1]:
            gv = conv_forward(x, c, None, stride, padding, dilation, groups, simplify)
        return gx, gv, None, None, None, None, None, None


class _UnfoldNdFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel_size, dilation, padding, stride):
        ctx.geometry = (x.shape[1], tuple(x.shape[2:]), kernel_size, dilation, padding, stride)
        return unfold(x, kernel_size, dilation, padding, stride)

    @staticmethod
    def backward(ctx, gu):
        channels, input_size, kernel_size, dilation, padding, stride = ctx.geometry
        return fold_autograd(gu, channels, input_size, kernel_size, dilation, padding, stride), None, None, None, None


class _FoldFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gu, channels, input_size, kernel_size, dilation, padding, stride):
        ctx.geometry = (kernel_size, dilation, padding, stride)
        return fold(gu, channels, input_size, kernel_size, dilation, padding, stride)

    @staticmethod
    def backward(ctx, c):
        kernel_size, dilation, padding, stride = ctx.geometry
        return unfold_autograd(c.contiguous(), kernel_size, dilation, padding, stride), None, None, None, None, None, None


def _factor_scale(op: int, x: Tensor, out: Sequence[int], scale) -> float:
    """The scalar the reference multiplies with (``convNd_kfc.py:105``, ``convNd_kfac_reduce.py:106-108``),
    rounded through ``x.dtype`` like its ``Tensor([..]).to(x.dtype)``."""
    if scale is not None:
        return float(scale)
    B = x.shape[0]
    denom = B
    if op != nat.OP_KFC:
        o_total = 1
        for o in out:
            o_total *= o
        denom = B * o_total**2
    return float(torch.tensor([1.0 / denom]).to(x.dtype)) if denom > 0 else 0.0


class _FactorFunction(torch.autograd.Function):
    """KFC / KFAC-reduce with the backward of the definition: unfold -> outer products -> fold."""

    @staticmethod
    def forward(ctx, x, op, kernel_size, stride, padding, dilation, groups, simplify, scale):
        ctx.save_for_backward(x)
        _, _, _, _, _, out = resolve_geometry(tuple(x.shape[2:]), kernel_size, stride, padding, dilation)
        ctx.scale = _factor_scale(op, x, out, scale)
        ctx.hyper = (op, kernel_size, stride, padding, dilation, groups)
        return _factor(op, x, kernel_size, stride, padding, dilation, groups, simplify, ctx.scale)

    @staticmethod
    def backward(ctx, c):
        (x,) = ctx.saved_tensors
        op, kernel_size, stride, padding, dilation, groups = ctx.hyper
        B, channels = x.shape[0], x.shape[1]
        u = unfold_autograd(x, kernel_size, dilation, padding, stride)  # [B, G*A, O]
        A, O = u.shape[1] // groups, u.shape[2]
        u = u.reshape(B, groups, A, O)
        sym = (c + c.transpose(1, 2)) * ctx.scale
        if op == nat.OP_KFC:
            gu = torch.einsum("gij,ngjo->ngio", sym, u)
        else:
            gm = torch.einsum("gij,ngj->ngi", sym, u.sum(-1))
            gu = gm.unsqueeze(-1).expand(B, groups, A, O)
        gx = fold_autograd(gu.reshape(B, groups * A, O).contiguous(), channels, tuple(x.shape[2:]), kernel_size,
                           dilation, padding, stride)
        return gx, None, None, None, None, None, None, None, None


class _UnfoldTransposeFunction(torch.autograd.Function):
    """Backward of the transpose-convolution unfolding: ``gx[b,c,o] = sum_k gu[b,(c,k), s o + d k - p]`` — a
    grouped convolution of ``gu`` (``C*K -> C`` channels, ``groups = C``) with a one-hot kernel."""

    @staticmethod
    def forward(ctx, x, kernel_size, stride, padding, output_padding, dilation):
        ctx.hyper = (kernel_size, stride, padding, dilation)
        ctx.x_shape = tuple(x.shape)
        return unfold_transpose(x, kernel_size, stride, padding, output_padding, dilation)

    @staticmethod
    def backward(ctx, gu):
        kernel_size, stride, padding, dilation = ctx.hyper
        B, C = ctx.x_shape[:2]
        N = len(ctx.x_shape) - 2
        t_k = _tuple(kernel_size, N)
        k_total = 1
        for k in t_k:
            k_total *= k
        # spatial sizes of gu's last axis: the associated convolution's input sizes
        t_s, t_p, t_d = _tuple(stride, N), _tuple(padding, N), _tuple(dilation, N)
        total = gu.shape[2]
        base = [get_conv_input_size(o, k, s, p, 0, d) for o, k, s, p, d in zip(ctx.x_shape[2:], t_k, t_s, t_p, t_d)]
        # output_padding only enlarges the input sizes; recover it from the element count dimension by dimension
        in_sizes = _recover_sizes(base, t_s, total)
        onehot = torch.eye(k_total, dtype=gu.dtype, device=gu.device).reshape(1, k_total, *t_k)
        onehot = onehot.expand(C, k_total, *t_k).contiguous()
        gx = conv_forward(gu.reshape(B, C * k_total, *in_sizes), onehot, None, stride, padding, dilation, C, True)
        return gx, None, None, None, None, None


def _recover_sizes(base: Sequence[int], strides: Sequence[int], total: int) -> Tuple[int, ...]:
    """Input sizes ``base_d + output_padding_d`` (``0 <= output_padding_d < stride_d``) whose product is ``total``."""
    def search(d, rest):
        if d == len(base):
            return () if rest == 1 else None
        for extra in range(max(1, strides[d])):
            size = base[d] + extra
            if size > 0 and rest % size == 0:
                tail = search(d + 1, rest // size)
                if tail is not None:
                    return (size,) + tail
        return None

    sizes = search(0, total)
    if sizes is None:
        raise RuntimeError("unfoldNd_transpose backward: cannot recover the unfolded sizes")
    return sizes


def conv_forward(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, simplify=True) -> Tensor:
    """Differentiable N-d convolution on the native kernels (``functionals/conv.py:11-69``)."""
    _conv_check(x, weight, bias, groups)
    if _needs_grad(x, weight, bias):
        return _ConvNdFunction.apply(x, weight, bias, stride, padding, dilation, groups, simplify)
    return conv_forward_raw(x, weight, bias, stride, padding, dilation, groups, simplify)


def conv_input_vjp(weight, v, input_size, stride=1, padding=0, dilation=1, groups=1, simplify=True) -> Tensor:
    """Differentiable input VJP (network of ``convNd_input_vjp.py:53-60``)."""
    if _needs_grad(weight, v):
        return _InputVjpFunction.apply(weight, v, input_size, stride, padding, dilation, groups, simplify)
    return conv_input_vjp_raw(weight, v, input_size, stride, padding, dilation, groups, simplify)


def conv_weight_vjp(x, v, kernel_size, dilation=1, padding=0, stride=1, groups=1, simplify=True) -> Tensor:
    """Differentiable weight VJP (network of ``convNd_weight_vjp.py:51-58``)."""
    if _needs_grad(x, v):
        return _WeightVjpFunction.apply(x, v, kernel_size, dilation, padding, stride, groups, simplify)
    return conv_weight_vjp_raw(x, v, kernel_size, dilation, padding, stride, groups, simplify)


def unfold_autograd(x, kernel_size, dilation=1, padding=0, stride=1) -> Tensor:
    if _needs_grad(x):
        return _UnfoldNdFunction.apply(x, kernel_size, dilation, padding, stride)
    return unfold(x, kernel_size, dilation, padding, stride)


def fold_autograd(gu, channels, input_size, kernel_size, dilation=1, padding=0, stride=1) -> Tensor:
    if _needs_grad(gu):
        return _FoldFunction.apply(gu, channels, input_size, kernel_size, dilation, padding, stride)
    return fold(gu, channels, input_size, kernel_size, dilation, padding, stride)


def unfold_transpose_autograd(x, kernel_size, stride=1, padding=0, output_padding=0, dilation=1) -> Tensor:
    if _needs_grad(x):
        return _UnfoldTransposeFunction.apply(x, kernel_size, stride, padding, output_padding, dilation)
    return unfold_transpose(x, kernel_size, stride, padding, output_padding, dilation)


# --------------------------------------------------------------------------------------------
# torch.einsum interception
# --------------------------------------------------------------------------------------------
@dataclass
class Plan:
    """What an expression computes, in terms the native evaluator understands."""

    kind: str                       # forward | input_vjp | weight_vjp | unfold | kfc | kfac_reduce
    tensors: Dict[str, Tensor]      # the un-rearranged tensors (x, weight, v)
    hyper: Dict[str, Any]           # stride / padding / dilation / groups / kernel_size / input_size
    shape: Tuple[int, ...]          # shape the reference reshapes the einsum result to
    simplify: bool
    operands: List[Any] = field(default_factory=list)  # exactly the operand objects returned

    def run(self) -> Tensor:
        t, h = self.tensors, self.hyper
        if self.kind == "forward":
            return conv_forward(t["x"], t["weight"], None, h["stride"], h["padding"], h["dilation"],
                                h["groups"], self.simplify)
        if self.kind == "input_vjp":
            return conv_input_vjp(t["weight"], t["v"], h["input_size"], h["stride"], h["padding"],
                                  h["dilation"], h["groups"], self.simplify)
        if self.kind == "weight_vjp":
            return conv_weight_vjp(t["x"], t["v"], h["kernel_size"], h["dilation"], h["padding"],
                                   h["stride"], h["groups"], self.simplify)
        if self.kind == "unfold":
            return unfold_autograd(t["x"], h["kernel_size"], h["dilation"], h["padding"], h["stride"])
        if self.kind == "kfc":
            return kfc(t["x"], h["kernel_size"], h["stride"], h["padding"], h["dilation"], h["groups"],
                       self.simplify)
        if self.kind == "kfac_reduce":
            return kfac_reduce(t["x"], h["kernel_size"], h["stride"], h["padding"], h["dilation"],
                               h["groups"], self.simplify)
        if self.kind == "unfold_transpose":
            return unfold_transpose_autograd(t["x"], h["kernel_size"], h["stride"], h["padding"],
                                             h["output_padding"], h["dilation"])
        if self.kind == "kfac_reduce_transpose":
            return kfac_reduce_transpose(t["x"], h["kernel_size"], h["stride"], h["padding"],
                                         h["output_padding"], h["dilation"], h["groups"], self.simplify)
        raise AssertionError(self.kind)


class Equation(str):
    """An einsum equation string that remembers which convolution network it encodes."""

    plan: Optional[Plan] = None

    def __new__(cls, text: str, plan: Optional[Plan] = None):
        self = super().__new__(cls, text)
        self.plan = plan
        return self


class PlannedOperand(Tensor):
    """Marker type for the first operand of a planned expression that has no index pattern left.

    When the simplifier removes every pattern (dense / 1x1 / down-sampling cases,
    ``simplifications/opt.py:196-300``) nothing in the operand list would route ``torch.einsum`` to
    the native evaluator, so the first data operand is tagged with this zero-cost subclass (same
    storage, same autograd history).  Every other torch function sees a plain tensor.
    """

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.einsum or func is torch.functional.einsum:
            return einsum(*args, **kwargs)
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)


def tag_for_dispatch(operands: List[Any]) -> List[Any]:
    """Make sure ``torch.einsum`` over ``operands`` reaches :func:`einsum` (see PlannedOperand)."""
    from einconv_b200.conv_index_pattern import IndexPattern

    if any(isinstance(op, IndexPattern) for op in operands):
        return operands
    for pos, op in enumerate(operands):
        if isinstance(op, Tensor):
            operands = list(operands)
            operands[pos] = op.as_subclass(PlannedOperand)
            break
    return operands


def plain(operands: Sequence[Any]) -> List[Tensor]:
    """Ordinary dense tensors for ``operands``: patterns written out, dispatch tags removed."""
    return [_dense(op) for op in operands]


def _einsum_output_shape(equation: str, operands: Sequence[Any]) -> Tuple[int, ...]:
    lhs, rhs = equation.replace(" ", "").split("->")
    dims: Dict[str, int] = {}
    for letters, op in zip(lhs.split(","), operands):
        for letter, size in zip(letters, op.shape):
            dims[letter] = int(size)
    return tuple(dims[letter] for letter in rhs)


def _dense(op):
    from einconv_b200.conv_index_pattern import IndexPattern

    if isinstance(op, IndexPattern):
        return op.materialize()
    if isinstance(op, PlannedOperand):
        return op.as_subclass(Tensor)
    return op


def einsum(equation, *operands):
    """``torch.einsum`` for expressions that contain lazy index patterns.

    A planned expression (equation and operand list exactly as returned by one of this package's
    ``einsum_expression`` functions, all on CUDA) runs the native sm_100a kernel and comes back in
    the shape ``torch.einsum`` would have produced.  Anything else is a user-composed network: the
    patterns are written out and torch contracts them — on CUDA tensors only.
    """
    if len(operands) == 1 and isinstance(operands[0], (list, tuple)):
        operands = tuple(operands[0])
    tensors = [op for op in operands if isinstance(op, Tensor)]
    nat.require_cuda(*tensors)

    plan = getattr(equation, "plan", None)
    if (
        plan is not None
        and len(plan.operands) == len(operands)
        and all(a is b for a, b in zip(plan.operands, operands))
    ):
        result = plan.run()
        return result.reshape(_einsum_output_shape(str(equation), operands))

    return _pattern_aware_einsum(str(equation), list(operands))


# --------------------------------------------------------------------------------------------
# user-composed networks: index patterns as native gathers, never as one-hot GEMMs
# --------------------------------------------------------------------------------------------
def _apply_pattern(tensor: Tensor, axis: int, pattern, transpose: bool) -> Optional[Tensor]:
    """Contract ``tensor``'s ``axis`` with an index pattern ``Pi[k, o, i]`` without forming ``Pi``.

    ``transpose=False``: the axis is ``i``; result axes ``(k, o)`` hold ``tensor[i = s o + d k - p]`` (0 outside):
    the 1-d ``unfoldNd`` gather.  ``transpose=True``: the axis is ``o``; result axes ``(k, i)`` hold
    ``tensor[o]`` where ``i = s o + d k - p``: the transpose-convolution gather.  Both run the bit-exact native
    kernels and are differentiable.  Returns None when the pattern cannot be expressed (asymmetric padding in
    the transpose direction)."""
    h = pattern._pattern_hyperparams
    K, S, D, P = h["kernel_size"], h["stride"], h["dilation"], h["padding"]
    moved = tensor.movedim(axis, -1)
    rest = moved.shape[:-1]
    flat = moved.reshape(-1, 1, moved.shape[-1])
    if not transpose:
        out = unfold_autograd(flat, K, dilation=D, padding=P, stride=S)          # [R, K, O]
    else:
        from einconv_b200.utils import get_conv_paddings

        p_left, p_right = get_conv_paddings(K, S, P, D)
        if p_left != p_right:
            return None
        base = get_conv_input_size(flat.shape[-1], K, S, p_left, 0, D)
        extra = h["input_size"] - base
        if extra < 0 or extra >= max(S, D):
            return None
        out = unfold_transpose_autograd(flat, K, stride=S, padding=p_left, output_padding=extra, dilation=D)
    out = out.reshape(*rest, out.shape[-2], out.shape[-1])
    n = out.dim()
    return out.movedim((n - 2, n - 1), (axis, axis + 1))


def _pattern_aware_einsum(equation: str, operands: List[Any]) -> Tensor:
    """``torch.einsum`` for networks that mix lazy index patterns with ordinary tensors.

    The reference evaluates such a network as dense GEMMs against the one-hot ``Pi`` tensors
    (``conv_index_pattern.py:12-85``).  Here every pattern whose input index ``i`` (or, in the transpose
    direction, output index ``o``) is shared with exactly one data operand and is summed out is *applied* to
    that operand as an indexed gather (the native 1-d unfold kernels): the operand trades its ``i`` axis for a
    ``(k, o)`` pair and the pattern disappears from the network.  What is left is an ordinary contraction of
    data tensors.  Patterns that do not fit (their index survives into the output, is shared by several
    operands, ...) are written out densely like before."""
    from einconv_b200.conv_index_pattern import IndexPattern

    eq = equation.replace(" ", "")
    if "->" not in eq or "." in eq:
        with torch._C.DisableTorchFunctionSubclass():
            return torch.einsum(equation, *[_dense(op) for op in operands])
    lhs, rhs = eq.split("->")
    terms = lhs.split(",")
    ops = [op.as_subclass(Tensor) if isinstance(op, PlannedOperand) else op for op in operands]
    if len(terms) != len(ops):
        with torch._C.DisableTorchFunctionSubclass():
            return torch.einsum(equation, *[_dense(op) for op in operands])
    changed = True
    while changed:
        changed = False
        for pos, (term, op) in enumerate(zip(terms, ops)):
            if not isinstance(op, IndexPattern) or len(term) != 3 or len(set(term)) != 3:
                continue
            k, o, i = term
            for letter, transpose in ((i, False), (o, True)):
                others = [q for q, t in enumerate(terms) if q != pos and letter in t]
                if letter in rhs or len(others) != 1:
                    continue
                q = others[0]
                t = terms[q]
                new_pair = (k + o) if not transpose else (k + i)
                if isinstance(ops[q], IndexPattern) or t.count(letter) != 1 or any(c in t for c in new_pair):
                    continue
                applied = _apply_pattern(ops[q], t.index(letter), op, transpose)
                if applied is None:
                    continue
                ops[q] = applied
                terms[q] = t.replace(letter, new_pair)
                del ops[pos], terms[pos]
                changed = True
                break
            if changed:
                break
    with torch._C.DisableTorchFunctionSubclass():
        return torch.einsum(",".join(terms) + "->" + rhs, *[_dense(op) for op in ops])


def evaluate(equation, operands, shape=None) -> Tensor:
    """``torch.einsum(equation, *operands)`` (optionally ``.reshape(shape)``) on the native kernels.

    Also covers simplified expressions in which no index pattern is left (dense / 1x1 cases),
    which ``torch.einsum`` alone would hand to cuBLAS without this package noticing.
    """
    plan = getattr(equation, "plan", None)
    if (
        plan is not None
        and len(plan.operands) == len(operands)
        and all(a is b for a, b in zip(plan.operands, operands))
    ):
        nat.require_cuda(*[op for op in operands if isinstance(op, Tensor)])
        result = plan.run()
        return result.reshape(shape if shape is not None else _einsum_output_shape(str(equation), operands))
    result = einsum(equation, *operands)
    return result if shape is None else result.reshape(shape)
