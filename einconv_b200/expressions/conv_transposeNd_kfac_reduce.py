"""Einsum expression of the input-based KFAC-reduce factor of a transpose convolution.

Contract: ``einconv/expressions/conv_transposeNd_kfac_reduce.py:21-173``.  With ``[[X_b]]_T`` the
unfolded transpose-convolution input and ``u_b`` its row sums over the associated convolution's
input positions, the factor is ``1 / (B * prod(I)^2) * sum_b u_b u_b^T``.  Hyper-parameters are those
of the associated convolution; ``output_size`` (of the transpose convolution) may replace
``output_padding`` (``:113-127``).
"""

from math import prod
from typing import List, Optional, Tuple, Union

from einops import rearrange
from torch import Tensor

from einconv_b200.expressions._common import finish, letters, pattern_strs
from einconv_b200.expressions.utils import create_conv_index_patterns
from einconv_b200.utils import _tuple, get_conv_input_size


def einsum_expression(
    x: Tensor,
    kernel_size: Union[int, Tuple[int, ...]],
    stride: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, Tuple[int, ...]] = 0,
    output_padding: Union[int, Tuple[int, ...]] = 0,
    output_size: Optional[Union[int, Tuple[int, ...]]] = None,
    dilation: Union[int, Tuple[int, ...]] = 1,
    groups: int = 1,
    simplify: bool = True,
) -> Tuple[str, List[Tensor], Tuple[int, ...]]:
    """Equation, operands ``[x, *patterns, *patterns, x, scale]`` and output shape ``[groups, A, A]``."""
    N = x.dim() - 2
    lhs = [
        "n g c_out " + letters("o", N),
        *pattern_strs(N),
        *pattern_strs(N, k_suffix="_", o_suffix="_", i_suffix="_"),
        "n g c_out_ " + letters("o", N, "_"),
        "s",
    ]
    rhs = "g c_out " + letters("k", N) + " c_out_ " + letters("k", N, "_")

    conv_output_size = x.shape[2:]
    t_kernel_size, t_stride = _tuple(kernel_size, N), _tuple(stride, N)
    t_padding, t_dilation = _tuple(padding, N), _tuple(dilation, N)
    if output_size is not None:
        t_output_size = _tuple(output_size, N)
        t_output_padding = tuple(
            size - get_conv_input_size(out, K, S, P, 0, D)
            for size, out, K, S, P, D in zip(
                t_output_size, conv_output_size, t_kernel_size, t_stride, t_padding, t_dilation
            )
        )
    else:
        t_output_padding = _tuple(output_padding, N)
    conv_input_size = tuple(
        get_conv_input_size(out, K, S, P, OP, D)
        for out, K, S, P, OP, D in zip(
            conv_output_size, t_kernel_size, t_stride, t_padding, t_output_padding, t_dilation
        )
    )
    patterns = create_conv_index_patterns(
        N,
        input_size=conv_input_size,
        kernel_size=t_kernel_size,
        stride=t_stride,
        padding=t_padding,
        dilation=t_dilation,
        device=x.device,
        dtype=x.dtype,
    )
    x_ungrouped = rearrange(x, "n (g c_in) ... -> n g c_in ...", g=groups)
    batch_size, out_channels = x.shape[:2]
    scale = Tensor([1.0 / (batch_size * prod(conv_input_size) ** 2)]).to(x.device, x.dtype)
    operands = [x_ungrouped, *patterns, *patterns, x_ungrouped, scale]

    A = out_channels // groups * prod(t_kernel_size)
    shape = (groups, A, A)
    hyper = dict(kernel_size=t_kernel_size, stride=t_stride, padding=t_padding,
                 output_padding=t_output_padding, dilation=t_dilation, groups=groups)
    return finish("kfac_reduce_transpose", lhs, rhs, operands, shape, dict(x=x), hyper, simplify)
