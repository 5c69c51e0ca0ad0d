"""``nn.Conv{1,2,3}d`` for any number of spatial dimensions, on the sm_100a kernels.

Contract: ``einconv/modules/conv.py:16-206`` — constructor signature and checks (``:16-112``),
``from_nn_Conv`` (``:114-145``), PyTorch's parameter initialisation (``:147-158``), parameter
names ``weight`` / ``bias`` so checkpoints interchange with ``torch.nn`` layers.
"""

from __future__ import annotations

from math import sqrt
from typing import Tuple, Union

import torch
from torch import Tensor
from torch.nn import Conv1d, Conv2d, Conv3d, Module, Parameter, init

from einconv_b200.functionals import convNd
from einconv_b200.utils import _tuple, sync_parameters


class ConvNd(Module):
    """N-dimensional convolution layer; trainable through the native VJP kernels."""

    def __init__(
        self,
        N: int,
        in_channels: int,
        out_channels: int,
        kernel_size: Union[int, Tuple[int, ...]],
        stride: Union[int, Tuple[int, ...]] = 1,
        padding: Union[int, Tuple[int, ...], str] = 0,
        dilation: Union[int, Tuple[int, ...]] = 1,
        groups: int = 1,
        bias: bool = True,
        padding_mode: str = "zeros",
        device: Union[None, torch.device] = None,
        dtype: Union[None, torch.dtype] = None,
        simplify: bool = True,
    ) -> None:
        """Create the layer; parameters follow PyTorch's convolution initialisation.

        Raises:
            NotImplementedError: For padding modes other than ``'zeros'``.
            ValueError: If ``groups`` is not positive or does not divide the channel counts.
        """
        super().__init__()
        if padding_mode != "zeros":
            raise NotImplementedError("Only padding_mode='zeros' supported.")
        if groups <= 0:
            raise ValueError("groups must be a positive integer")
        for name, channels in (("in_channels", in_channels), ("out_channels", out_channels)):
            if channels % groups != 0:
                raise ValueError(f"groups ({groups}) must divide {name} ({channels}).")

        self.N = N
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.simplify = simplify
        self.kernel_size = _tuple(kernel_size, N)
        self.stride = _tuple(stride, N)
        self.padding = padding if isinstance(padding, str) else _tuple(padding, N)
        self.padding_mode = padding_mode
        self.dilation = _tuple(dilation, N)
        self.groups = groups

        # None passes through to torch.empty like in the reference and torch.nn.Conv: torch.set_default_dtype,
        # set_default_device and meta / device contexts then apply (modules/conv.py:108 of the reference)
        factory = dict(device=device, dtype=dtype)
        self.weight = Parameter(
            torch.empty((out_channels, in_channels // groups, *self.kernel_size), **factory)
        )
        if bias:
            self.bias = Parameter(torch.empty((out_channels,), **factory))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    @classmethod
    def from_nn_Conv(
        cls, conv_module: Union[Conv1d, Conv2d, Conv3d], simplify: bool = True
    ) -> ConvNd:
        """Build a ``ConvNd`` with the hyper-parameters and parameter values of a torch layer."""
        N = {Conv1d: 1, Conv2d: 2, Conv3d: 3}[conv_module.__class__]
        layer = cls(
            N,
            conv_module.in_channels,
            conv_module.out_channels,
            conv_module.kernel_size,
            stride=conv_module.stride,
            padding=conv_module.padding,
            dilation=conv_module.dilation,
            groups=conv_module.groups,
            bias=conv_module.bias is not None,
            padding_mode=conv_module.padding_mode,
            device=conv_module.weight.device,
            dtype=conv_module.weight.dtype,
            simplify=simplify,
        )
        sync_parameters(conv_module, layer)
        return layer

    def reset_parameters(self) -> None:
        """Kaiming-uniform(a=sqrt(5)) weight, uniform(+-1/sqrt(fan_in)) bias, as in torch."""
        init.kaiming_uniform_(self.weight, a=sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            if fan_in != 0:
                bound = 1 / sqrt(fan_in)
                init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        """``[batch_size, in_channels, *input_sizes] -> [batch_size, out_channels, *output_sizes]``."""
        return convNd(
            x,
            self.weight,
            bias=self.bias,
            stride=self.stride,
            padding=self.padding,
            dilation=self.dilation,
            groups=self.groups,
            simplify=self.simplify,
        )

    def extra_repr(self) -> str:
        """Arguments that differ from their defaults, in torch's style."""
        parts = [
            f"{self.N}, {self.in_channels}, {self.out_channels}",
            f"kernel_size={self.kernel_size}",
            f"stride={self.stride}",
        ]
        if isinstance(self.padding, str) or any(p != 0 for p in self.padding):
            parts.append(f"padding={self.padding}")
        if any(d != 1 for d in self.dilation):
            parts.append(f"dilation={self.dilation}")
        if self.groups != 1:
            parts.append(f"groups={self.groups}")
        if self.bias is None:
            parts.append("bias=False")
        if self.padding_mode != "zeros":
            parts.append(f"padding_mode={self.padding_mode}")
        if self.simplify is False:
            parts.append("simplify=False")
        return ", ".join(parts)
