"""Einsum expression of input unfolding (im2col).

Contract: ``einconv/expressions/convNd_unfold.py:10-79`` — network
``n c_in i.. , {k_d o_d i_d} -> n c_in k.. o..`` reshaped to ``(N, C_in * prod(K), prod(O))``.
"""

from math import prod
from typing import List, Tuple, Union

from torch import Tensor

from einconv_b200.expressions._common import finish, letters, pattern_strs
from einconv_b200.expressions.utils import create_conv_index_patterns
from einconv_b200.utils import _tuple


def einsum_expression(
    x: Tensor,
    kernel_size: Union[int, Tuple[int, ...]],
    stride: Union[int, Tuple[int, ...]] = 1,
    padding: Union[str, int, Tuple[int, ...]] = 0,
    dilation: Union[int, Tuple[int, ...]] = 1,
    simplify: bool = True,
) -> Tuple[str, List[Tensor], Tuple[int, ...]]:
    """Equation, operands ``[x, *patterns]`` and output shape
    ``[batch_size, in_channels * tot_kernel_size, tot_output_size]`` of the unfolded input."""
    N = x.dim() - 2
    lhs = ["n c_in " + letters("i", N), *pattern_strs(N)]
    rhs = "n c_in " + letters("k", N) + " " + letters("o", N)

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
    operands = [x, *patterns]
    t_kernel_size = _tuple(kernel_size, N)
    shape = (
        x.shape[0],
        x.shape[1] * prod(t_kernel_size),
        prod(int(p.shape[1]) for p in patterns),
    )
    hyper = dict(kernel_size=t_kernel_size, stride=stride, padding=padding, dilation=dilation)
    return finish("unfold", lhs, rhs, operands, shape, dict(x=x), hyper, simplify)
