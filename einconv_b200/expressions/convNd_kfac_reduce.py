"""Einsum expression of the input-based KFAC-reduce factor of a convolution.

Contract: ``einconv/expressions/convNd_kfac_reduce.py:21-124``.  With ``[[X_b]]`` the unfolded
input, ``u_b = [[X_b]] 1`` its row sums over output positions, the factor is
``Omega^ = 1 / (B * prod(O)^2) * sum_b u_b u_b^T``; as a network it differs from KFC only in that
the two pattern copies have independent output indices ``o_d`` and ``o_d_``.
"""

from math import prod
from typing import List, Tuple, Union

from einops import rearrange
from torch import Tensor

from einconv_b200.expressions._common import finish, letters, pattern_strs
from einconv_b200.expressions.utils import create_conv_index_patterns
from einconv_b200.utils import _tuple


def einsum_expression(
    x: Tensor,
    kernel_size: Union[int, Tuple[int, ...]],
    stride: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, str, Tuple[int, ...]] = 0,
    dilation: Union[int, Tuple[int, ...]] = 1,
    groups: int = 1,
    simplify: bool = True,
) -> Tuple[str, List[Tensor], Tuple[int, ...]]:
    """Equation, operands ``[x, *patterns, *patterns, x, scale]`` and output shape
    ``[groups, A, A]`` of the KFAC-reduce factor."""
    N = x.dim() - 2
    lhs = [
        "n g c_in " + letters("i", N),
        *pattern_strs(N),
        *pattern_strs(N, k_suffix="_", o_suffix="_", i_suffix="_"),
        "n g c_in_ " + letters("i", N, "_"),
        "s",
    ]
    rhs = "g c_in " + letters("k", N) + " c_in_ " + letters("k", N, "_")

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
    x_ungrouped = rearrange(x, "n (g c_in) ... -> n g c_in ...", g=groups)
    output_tot_size = prod(int(p.shape[1]) for p in patterns)
    batch_size = x.shape[0]
    scale = Tensor([1.0 / (batch_size * output_tot_size**2)]).to(x.device, x.dtype)
    operands = [x_ungrouped, *patterns, *patterns, x_ungrouped, scale]

    t_kernel_size = _tuple(kernel_size, N)
    A = x.shape[1] // groups * prod(t_kernel_size)
    shape = (groups, A, A)
    hyper = dict(kernel_size=t_kernel_size, stride=stride, padding=padding, dilation=dilation,
                 groups=groups)
    return finish("kfac_reduce", lhs, rhs, operands, shape, dict(x=x), hyper, simplify)
