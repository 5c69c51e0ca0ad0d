"""Index-pattern operands of convolution tensor networks, without materialising them.

Reference: ``einconv/conv_index_pattern.py:12-85`` (``index_pattern``, built with
``conv1d`` + ``scatter_add_``) and ``:88-146`` (``index_pattern_logical``, the closed
form).  The pattern of one spatial dimension is the one-hot tensor

    Pi[k, o, i] = 1  iff  i == stride * o + dilation * k - padding_left  and 0 <= i < I

Here ``index_pattern`` returns an :class:`IndexPattern`: a ``torch.Tensor`` subclass
that only *describes* that tensor (shape ``[K, O, I]``, dtype, device and the
``_pattern_hyperparams`` dict the reference attaches, ``conv_index_pattern.py:77-83``).
The sm_100a kernels evaluate the index law in registers, so on the hot path the dense
one-hot tensor never exists.  It is only written out if somebody asks for its values
(any torch op other than the ``torch.einsum`` calls the native evaluator recognises).
"""

from typing import Union

import torch
from torch import Tensor
from torch.utils._pytree import tree_map

from einconv_b200.utils import cpu, get_conv_output_size, get_conv_paddings


def _dense_pattern(
    input_size: int,
    kernel_size: int,
    stride: int,
    padding: Union[int, str],
    dilation: int,
    device: torch.device,
    dtype: torch.dtype,
) -> Tensor:
    """Write out Pi[k, o, i] from the closed-form index law (vectorised)."""
    output_size = get_conv_output_size(input_size, kernel_size, stride, padding, dilation)
    p_left, _ = get_conv_paddings(kernel_size, stride, padding, dilation)
    k = torch.arange(kernel_size, device=device).view(-1, 1, 1)
    o = torch.arange(max(output_size, 0), device=device).view(1, -1, 1)
    i = torch.arange(input_size, device=device).view(1, 1, -1)
    return (i == stride * o + dilation * k - p_left).to(dtype)


class IndexPattern(Tensor):
    """Lazy one-hot connectivity tensor ``[kernel_size, output_size, input_size]``.

    Behaves like the dense tensor the reference returns: ``shape``/``dtype``/``device``
    are available without any device work, ``_pattern_hyperparams`` carries the
    geometry (the attribute the simplifier keys on, ``simplifications/opt.py:200``),
    and every torch operation sees the dense values (written on first use and cached).
    ``torch.einsum`` calls whose operands include an ``IndexPattern`` are routed to
    :func:`einconv_b200.evaluator.einsum`, which runs the sm_100a kernels when it
    recognises the network.
    """

    @staticmethod
    def __new__(cls, input_size, kernel_size, stride, padding, dilation, device, dtype):
        output_size = get_conv_output_size(
            input_size, kernel_size, stride, padding, dilation
        )
        self = Tensor._make_wrapper_subclass(
            cls,
            (kernel_size, max(output_size, 0), input_size),
            dtype=dtype,
            device=device,
            requires_grad=False,
        )
        self._pattern_hyperparams = {
            "input_size": input_size,
            "kernel_size": kernel_size,
            "stride": stride,
            "padding": padding,
            "dilation": dilation,
        }
        self._dense = None
        return self

    # -- geometry ---------------------------------------------------------------------
    @property
    def padding_left(self) -> int:
        h = self._pattern_hyperparams
        return get_conv_paddings(h["kernel_size"], h["stride"], h["padding"], h["dilation"])[0]

    def materialize(self) -> Tensor:
        """Return the dense one-hot tensor (plain ``torch.Tensor``), cached."""
        if self._dense is None:
            h = self._pattern_hyperparams
            with torch._C.DisableTorchFunctionSubclass():
                dense = _dense_pattern(
                    h["input_size"],
                    h["kernel_size"],
                    h["stride"],
                    h["padding"],
                    h["dilation"],
                    self.device,
                    self.dtype,
                )
            self._dense = dense
        return self._dense

    def __repr__(self) -> str:  # noqa: D105
        h = self._pattern_hyperparams
        return (
            f"IndexPattern(shape={tuple(self.shape)}, dtype={self.dtype}, "
            f"device={self.device}, {h})"
        )

    # -- dispatch ---------------------------------------------------------------------
    _METADATA = None  # filled lazily: functions answered from the wrapper itself

    @classmethod
    def _metadata_funcs(cls):
        if cls._METADATA is None:
            cls._METADATA = {
                Tensor.shape.__get__,
                Tensor.dtype.__get__,
                Tensor.device.__get__,
                Tensor.ndim.__get__,
                Tensor.is_cuda.__get__,
                Tensor.layout.__get__,
                Tensor.requires_grad.__get__,
                Tensor.grad_fn.__get__,
                Tensor.size,
                Tensor.dim,
                Tensor.numel,
                Tensor.element_size,
                Tensor.is_floating_point,
                Tensor.is_complex,
                Tensor.__len__,
            }
        return cls._METADATA

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.einsum or func is torch.functional.einsum:
            from einconv_b200 import evaluator

            return evaluator.einsum(*args, **kwargs)
        if func in cls._metadata_funcs():
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)

        def dense(obj):
            return obj.materialize() if isinstance(obj, IndexPattern) else obj

        with torch._C.DisableTorchFunctionSubclass():
            return func(*tree_map(dense, args), **tree_map(dense, kwargs))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # Safety net: an aten op reached the storage-less wrapper directly.
        def dense(obj):
            return obj.materialize() if isinstance(obj, IndexPattern) else obj

        return func(*tree_map(dense, args), **tree_map(dense, kwargs or {}))


def index_pattern(
    input_size: int,
    kernel_size: int,
    stride: int = 1,
    padding: Union[int, str] = 0,
    dilation: int = 1,
    device: torch.device = cpu,
    dtype: torch.dtype = torch.bool,
) -> Tensor:
    """Connectivity pattern of a convolution along one dimension.

    Same signature and meaning as the reference's ``index_pattern``
    (``einconv/conv_index_pattern.py:12-85``): element ``[k, o, i]`` is true iff input
    ``i`` reaches output ``o`` through kernel tap ``k``.  The returned
    :class:`IndexPattern` is lazy; use ``.materialize()`` (or any torch op) for values.
    """
    device = torch.device(device) if not isinstance(device, torch.device) else device
    # validates the padding string / 'same'+stride combination like the reference does
    get_conv_paddings(kernel_size, stride, padding, dilation)
    return IndexPattern(input_size, kernel_size, stride, padding, dilation, device, dtype)


def index_pattern_logical(
    input_size: int,
    kernel_size: int,
    stride: int = 1,
    padding: Union[int, str] = 0,
    dilation: int = 1,
    device: torch.device = cpu,
    dtype: torch.dtype = torch.bool,
) -> Tensor:
    """Dense pattern from the closed-form law (``conv_index_pattern.py:88-146``).

    Returns a plain tensor with ``_pattern_hyperparams`` attached, like the reference.
    """
    pattern = _dense_pattern(
        input_size, kernel_size, stride, padding, dilation, torch.device(device), dtype
    )
    pattern._pattern_hyperparams = {
        "input_size": input_size,
        "kernel_size": kernel_size,
        "stride": stride,
        "padding": padding,
        "dilation": dilation,
    }
    return pattern
