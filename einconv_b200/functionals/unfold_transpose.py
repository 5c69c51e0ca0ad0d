"""Input unfolding of a transpose convolution for any number of spatial dimensions.

Contract: ``einconv/functionals/unfold_transpose.py:10-70``; the reference's einsum (``:70``) is
replaced by the bit-exact gather kernel ``einconv_unfold_transpose``.
"""

from typing import Tuple, Union

from torch import Tensor

from einconv_b200 import evaluator


def unfoldNd_transpose(
    x: Tensor,
    kernel_size: Union[int, Tuple[int, ...]],
    stride: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, Tuple[int, ...]] = 0,
    output_padding: Union[int, Tuple[int, ...]] = 0,
    dilation: Union[int, Tuple[int, ...]] = 1,
    simplify: bool = True,
) -> Tensor:
    """Unfold the input ``[batch_size, out_channels, *output_sizes]`` of a transpose convolution.

    All hyper-parameters are those of the associated convolution.  Returns
    ``[batch_size, out_channels * tot_kernel_size, tot_input_size]`` with index structure
    ``n (c_out k1 k2 ...) (i1 i2 ...)``.  ``simplify`` is accepted for API compatibility: the gather
    is one kernel for every geometry.
    """
    del simplify
    return evaluator.unfold_transpose_autograd(
        x, kernel_size, stride=stride, padding=padding, output_padding=output_padding, dilation=dilation
    )
