"""Einsum expression of an N-d convolution's forward pass.

Contract: ``einconv/expressions/convNd_forward.py:13-82`` — network
``n g c_in i.. , {k_d o_d i_d} , g c_out c_in k.. -> n g c_out o..`` with pattern dtype/device
taken from the weight and result shape ``(N, C_out, *O)``.
"""

from typing import List, Tuple, Union

from einops import rearrange
from torch import Tensor
from torch.nn import Parameter

from einconv_b200.expressions._common import finish, letters, pattern_strs
from einconv_b200.expressions.utils import create_conv_index_patterns


def einsum_expression(
    x: Tensor,
    weight: Union[Tensor, Parameter],
    stride: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, str, Tuple[int, ...]] = 0,
    dilation: Union[int, Tuple[int, ...]] = 1,
    groups: int = 1,
    simplify: bool = True,
) -> Tuple[str, List[Union[Tensor, Parameter]], Tuple[int, ...]]:
    """Equation, operands ``[x (ungrouped), *patterns, weight (ungrouped)]`` and output shape
    ``[batch_size, out_channels, *output_sizes]`` of a convolution's forward pass."""
    N = x.dim() - 2
    lhs = [
        "n g c_in " + letters("i", N),
        *pattern_strs(N),
        "g c_out c_in " + letters("k", N),
    ]
    rhs = "n g c_out " + letters("o", N)

    patterns = create_conv_index_patterns(
        N,
        input_size=x.shape[2:],
        kernel_size=weight.shape[2:],
        stride=stride,
        padding=padding,
        dilation=dilation,
        device=weight.device,
        dtype=weight.dtype,
    )
    operands = [
        rearrange(x, "n (g c_in) ... -> n g c_in ...", g=groups),
        *patterns,
        rearrange(weight, "(g c_out) ... -> g c_out ...", g=groups),
    ]
    shape = (x.shape[0], weight.shape[0], *[p.shape[1] for p in patterns])
    hyper = dict(stride=stride, padding=padding, dilation=dilation, groups=groups)
    return finish("forward", lhs, rhs, operands, shape, dict(x=x, weight=weight), hyper, simplify)
