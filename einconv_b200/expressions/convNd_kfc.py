"""Einsum expression of the input-based KFC factor of a convolution.

Contract: ``einconv/expressions/convNd_kfc.py:20-121``.  With ``[[X_b]]`` the unfolded input of
sample ``b`` (shape ``(C_in/G * prod(K)) x prod(O)`` per group), the factor is
``Omega = 1/B * sum_b [[X_b]] [[X_b]]^T``, written as the 7-operand network
``x, {Pi_d}, {Pi_d'}, x', s`` in which both pattern copies share the output index ``o_d``.
Result shape ``(G, A, A)`` with ``A = C_in/G * prod(K)``.
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
    ``[groups, A, A]`` of the KFC factor."""
    N = x.dim() - 2
    lhs = [
        "n g c_in " + letters("i", N),
        *pattern_strs(N),
        *pattern_strs(N, k_suffix="_", i_suffix="_"),  # shares o_d with the first copy
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
    batch_size = x.shape[0]
    scale = Tensor([1.0 / batch_size]).to(x.device).to(x.dtype)
    operands = [x_ungrouped, *patterns, *patterns, x_ungrouped, scale]

    t_kernel_size = _tuple(kernel_size, N)
    A = x.shape[1] // groups * prod(t_kernel_size)
    shape = (groups, A, A)
    hyper = dict(kernel_size=t_kernel_size, stride=stride, padding=padding, dilation=dilation,
                 groups=groups)
    return finish("kfc", lhs, rhs, operands, shape, dict(x=x), hyper, simplify)
