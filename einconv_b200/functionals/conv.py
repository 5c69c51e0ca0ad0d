"""N-d generalisation of ``torch.nn.functional.conv{1,2,3}d`` on the sm_100a kernels.

Contract: ``einconv/functionals/conv.py:11-69`` (signature, argument checks ``:72-116``, bias
``:63-67``).  Where the reference evaluates ``torch.einsum(equation, *operands).reshape(shape)``
(``:61``), this calls the native evaluator; no expression is built and no index pattern exists.
"""

from typing import Tuple, Union

from torch import Tensor
from torch.nn import Parameter

from einconv_b200 import evaluator


def convNd(
    x: Tensor,
    weight: Union[Tensor, Parameter],
    bias: Union[Tensor, Parameter, None] = None,
    stride: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, str, Tuple[int, ...]] = 0,
    dilation: Union[int, Tuple[int, ...]] = 1,
    groups: int = 1,
    simplify: bool = True,
) -> Tensor:
    """Convolution of ``x`` ``[batch_size, in_channels, *input_sizes]`` with ``weight``
    ``[out_channels, in_channels / groups, *kernel_size]`` for any number of spatial dimensions.

    Args:
        x: Convolution input.
        weight: Kernel of the convolution.
        bias: Optional bias of shape ``[out_channels]``. Default: ``None``.
        stride: Int or ``N``-tuple. Default: ``1``.
        padding: Int, ``N``-tuple, ``'same'`` or ``'valid'``. Default: ``0``.
        dilation: Int or ``N``-tuple. Default: ``1``.
        groups: Number of channel groups. Default: ``1``.
        simplify: ``True`` lets the dispatcher use specialised kernels (depthwise stencil,
            dense/1x1, tensor cores); ``False`` forces the single generic kernel. The results
            agree, as they do for the reference's simplified and unsimplified expressions.

    Returns:
        ``[batch_size, out_channels, *output_sizes]`` with index structure ``n (g c_out) o1 o2 ...``.

    Raises:
        ValueError: For inconsistent weight / bias / groups (same checks as the reference).
        RuntimeError: If a tensor is not on a CUDA device (there is no CPU path).
    """
    return evaluator.conv_forward(
        x,
        weight,
        bias=bias,
        stride=stride,
        padding=padding,
        dilation=dilation,
        groups=groups,
        simplify=simplify,
    )


def _check_args(x, weight, bias=None, groups: int = 1) -> None:
    """Argument checks of ``convNd`` (kept under the reference's name)."""
    evaluator._conv_check(x, weight, bias, groups)
