"""``nn.Unfold`` (im2col) for any number of spatial dimensions.

Contract: ``einconv/modules/unfold.py:11-70``.
"""

from typing import Tuple, Union

from torch import Tensor
from torch.nn import Module

from einconv_b200.functionals import unfoldNd


class UnfoldNd(Module):
    """Extracts sliding local blocks from a batched input with ``N`` spatial dimensions."""

    def __init__(
        self,
        kernel_size: Union[int, Tuple[int, ...]],
        dilation: Union[int, Tuple[int, ...]] = 1,
        padding: Union[int, Tuple[int, ...], str] = 0,
        stride: Union[int, Tuple[int, ...]] = 1,
        simplify: bool = True,
    ):
        """Store the hyper-parameters; ``N`` is taken from the input at call time."""
        super().__init__()
        self._kernel_size = kernel_size
        self._dilation = dilation
        self._padding = padding
        self._stride = stride
        self._simplify = simplify

    def forward(self, x: Tensor) -> Tensor:
        """``[batch_size, in_channels, *input_sizes] ->
        [batch_size, in_channels * tot_kernel_size, tot_output_size]``."""
        return unfoldNd(
            x,
            self._kernel_size,
            dilation=self._dilation,
            padding=self._padding,
            stride=self._stride,
            simplify=self._simplify,
        )
