"""Einsum expression of input unfolding for a transpose convolution.

Contract: ``einconv/expressions/conv_transposeNd_unfold.py:10-136`` — network
``n c_out o.. , {k_d o_d i_d} -> n c_out k.. i..`` reshaped to
``(N, C_out * prod(K), prod(I))``.  All hyper-parameters are those of the ASSOCIATED convolution,
whose input sizes ``I`` follow from ``x``'s spatial sizes and ``output_padding``
(``einconv/utils.py:163-205``).  The patterns are the same ``Pi[k, o, i]`` as for a convolution,
contracted over ``o`` instead of ``i``.
"""

from math import prod
from typing import List, Tuple, Union

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
    dilation: Union[int, Tuple[int, ...]] = 1,
    simplify: bool = True,
) -> Tuple[str, List[Tensor], Tuple[int, ...]]:
    """Equation, operands ``[x, *patterns]`` and output shape
    ``[batch_size, out_channels * tot_kernel_size, tot_input_size]``."""
    N = x.dim() - 2
    lhs = ["n c_out " + letters("o", N), *pattern_strs(N)]
    rhs = "n c_out " + letters("k", N) + " " + letters("i", N)

    t_kernel_size = _tuple(kernel_size, N)
    t_stride, t_padding = _tuple(stride, N), _tuple(padding, N)
    t_dilation, t_output_padding = _tuple(dilation, N), _tuple(output_padding, N)
    t_input_size = tuple(
        get_conv_input_size(o, k, s, p, op, d)
        for o, k, s, p, op, d in zip(
            x.shape[2:], t_kernel_size, t_stride, t_padding, t_output_padding, t_dilation
        )
    )
    patterns = create_conv_index_patterns(
        N,
        input_size=t_input_size,
        kernel_size=t_kernel_size,
        stride=t_stride,
        padding=t_padding,
        dilation=t_dilation,
        device=x.device,
        dtype=x.dtype,
    )
    operands = [x, *patterns]
    shape = (x.shape[0], x.shape[1] * prod(t_kernel_size), prod(t_input_size))
    hyper = dict(kernel_size=t_kernel_size, stride=t_stride, padding=t_padding,
                 output_padding=t_output_padding, dilation=t_dilation)
    return finish("unfold_transpose", lhs, rhs, operands, shape, dict(x=x), hyper, simplify)
