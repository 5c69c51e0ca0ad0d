"""Einsum expression of an N-d convolution's input VJP.

Contract: ``einconv/expressions/convNd_input_vjp.py:14-87`` — network
``n g c_out o.. , {k_d o_d i_d} , g c_out c_in k.. -> n g c_in i..``; the input size must be given
because strides make it ambiguous; result shape ``(N, G * C_in/G, *I)``.
"""

from typing import List, Tuple, Union

from einops import rearrange
from torch import Tensor
from torch.nn import Parameter

from einconv_b200.expressions._common import finish, letters, pattern_strs
from einconv_b200.expressions.utils import create_conv_index_patterns
from einconv_b200.utils import _tuple


def einsum_expression(
    weight: Union[Tensor, Parameter],
    v: Tensor,
    input_size: Union[int, Tuple[int, ...]],
    stride: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, str, Tuple[int, ...]] = 0,
    dilation: Union[int, Tuple[int, ...]] = 1,
    groups: int = 1,
    simplify: bool = True,
) -> Tuple[str, List[Union[Tensor, Parameter]], Tuple[int, ...]]:
    """Equation, operands ``[v (ungrouped), *patterns, weight (ungrouped)]`` and output shape
    ``[batch_size, in_channels, *input_sizes]`` of a convolution's input VJP."""
    N = weight.dim() - 2
    lhs = [
        "n g c_out " + letters("o", N),
        *pattern_strs(N),
        "g c_out c_in " + letters("k", N),
    ]
    rhs = "n g c_in " + letters("i", N)

    patterns = create_conv_index_patterns(
        N,
        input_size,
        kernel_size=weight.shape[2:],
        stride=stride,
        padding=padding,
        dilation=dilation,
        device=weight.device,
        dtype=weight.dtype,
    )
    operands = [
        rearrange(v, "n (g c_out) ... -> n g c_out ...", g=groups),
        *patterns,
        rearrange(weight, "(g c_out) ... -> g c_out ...", g=groups),
    ]
    t_input_size = _tuple(input_size, N)
    shape = (v.shape[0], groups * weight.shape[1], *t_input_size)
    hyper = dict(input_size=t_input_size, stride=stride, padding=padding, dilation=dilation,
                 groups=groups)
    return finish("input_vjp", lhs, rhs, operands, shape, dict(weight=weight, v=v), hyper, simplify)
