"""Tensor-free predicates of the simplifier, one spatial dimension at a time.

These are the conditions under which the reference rewrites a pattern
(``einconv/simplifications/opt.py:196-222`` down-sampling, ``:250-283`` dense; the
same predicates are stated on shapes only in ``simplifications/symbolic.py:518-544``).
The tensor-network rewrite in :mod:`einconv_b200.simplifications.opt` uses them, and so
does the native dispatcher, which maps them to kernel specialisations:

* ``dense`` with ``K == S == 1``  -> pointwise (plain GEMM) variants,
* ``dense`` with ``K == S > 1``   -> patchify: non-overlapping taps,
* ``downsampling`` (``S > K``)    -> input pixels that are never read.
"""

from typing import Union

from einconv_b200.utils import get_conv_paddings


def _symmetric_padding(kernel_size, stride, padding, dilation):
    """Integer padding, or ``None`` for a string padding with unequal sides."""
    if isinstance(padding, str):
        left, right = get_conv_paddings(kernel_size, stride, padding, dilation)
        return left if left == right else None
    return padding


def is_downsampling(
    input_size: int,
    kernel_size: int,
    stride: int,
    padding: Union[int, str],
    dilation: int,
) -> bool:
    """``S > K`` without padding/dilation and ``S | I``: some pixels are never read."""
    p = _symmetric_padding(kernel_size, stride, padding, dilation)
    if p is None:
        return False
    return stride > kernel_size and p == 0 and dilation == 1 and input_size % stride == 0


def is_dense(
    input_size: int,
    kernel_size: int,
    stride: int,
    padding: Union[int, str],
    dilation: int,
) -> bool:
    """``K == S`` without padding/dilation, tiling the input exactly."""
    p = _symmetric_padding(kernel_size, stride, padding, dilation)
    if p is None:
        return False
    span = kernel_size + (kernel_size - 1) * (dilation - 1)
    return (
        (input_size + 2 * p - span) % stride == 0
        and kernel_size == stride
        and p == 0
        and dilation == 1
    )
