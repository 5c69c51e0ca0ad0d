"""``torch.nn.functional.unfold`` (im2col) for any number of spatial dimensions.

Contract: ``einconv/functionals/unfold.py:10-56`` — note the argument order
``kernel_size, dilation, padding, stride`` (like ``F.unfold``).  The reference's einsum (``:56``)
is replaced by the bit-exact gather kernel.
"""

from typing import Tuple, Union

from torch import Tensor

from einconv_b200 import evaluator


def unfoldNd(
    x: Tensor,
    kernel_size: Union[int, Tuple[int, ...]],
    dilation: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, Tuple[int, ...], str] = 0,
    stride: Union[int, Tuple[int, ...]] = 1,
    simplify: bool = True,
) -> Tensor:
    """Extract sliding local blocks of ``x`` ``[batch_size, in_channels, *input_sizes]``.

    Returns ``[batch_size, in_channels * tot_kernel_size, tot_output_size]`` with index structure
    ``n (c_in k1 k2 ...) (o1 o2 ...)``.  ``simplify`` is accepted for API compatibility: the
    gather is one kernel for every geometry, so both settings run the same code.
    """
    del simplify
    return evaluator.unfold_autograd(
        x, kernel_size, dilation=dilation, padding=padding, stride=stride
    )
