"""Einsum expression of an N-d convolution's weight VJP.

Contract: ``einconv/expressions/convNd_weight_vjp.py:13-85`` — network
``n g c_in i.. , {k_d o_d i_d} , n g c_out o.. -> g c_out c_in k..``; note the keyword order
``kernel_size, dilation, padding, stride, groups``; result shape ``(C_out, C_in/G, *K)``.
"""

from typing import List, Tuple, Union

from einops import rearrange
from torch import Tensor

from einconv_b200.expressions._common import finish, letters, pattern_strs
from einconv_b200.expressions.utils import create_conv_index_patterns
from einconv_b200.utils import _tuple


def einsum_expression(
    x: Tensor,
    v: Tensor,
    kernel_size: Union[int, Tuple[int, ...]],
    dilation: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, str, Tuple[int, ...]] = 0,
    stride: Union[int, Tuple[int, ...]] = 1,
    groups: int = 1,
    simplify: bool = True,
) -> Tuple[str, List[Tensor], Tuple[int, ...]]:
    """Equation, operands ``[x (ungrouped), *patterns, v (ungrouped)]`` and output shape
    ``[out_channels, in_channels // groups, *kernel_size]`` of a convolution's weight VJP."""
    N = x.dim() - 2
    lhs = [
        "n g c_in " + letters("i", N),
        *pattern_strs(N),
        "n g c_out " + letters("o", N),
    ]
    rhs = "g c_out c_in " + letters("k", N)

    patterns = create_conv_index_patterns(
        N,
        input_size=x.shape[2:],
        kernel_size=kernel_size,
        stride=stride,
        padding=padding,
        dilation=dilation,
        device=x.device,
        dtype=x.dtype,
    )
    operands = [
        rearrange(x, "n (g c_in) ... -> n g c_in ...", g=groups),
        *patterns,
        rearrange(v, "n (g c_out) ... -> n g c_out ...", g=groups),
    ]
    t_kernel_size = _tuple(kernel_size, N)
    shape = (v.shape[1], x.shape[1] // groups, *t_kernel_size)
    hyper = dict(kernel_size=t_kernel_size, dilation=dilation, padding=padding, stride=stride,
                 groups=groups)
    return finish("weight_vjp", lhs, rhs, operands, shape, dict(x=x, v=v), hyper, simplify)
