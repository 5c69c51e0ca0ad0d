"""Convolution geometry helpers shared by the host code and the native kernels.

The formulas here are the contract of the device-side ``ConvGeom`` struct
(``include/einconv_b200.h``).  They restate the reference semantics of
``einconv/utils.py:12-48`` (``_tuple``), ``:97-132`` (``get_conv_paddings``),
``:135-160`` (``get_conv_output_size``), ``:163-205`` (``get_conv_input_size``)
and ``:51-76`` (``sync_parameters``); error types and messages follow the
reference so that callers catching them keep working.
"""

from typing import Any, List, Sequence, Tuple, Union

import torch
from torch.nn import Module

cpu = torch.device("cpu")

IntOrTuple = Union[int, Tuple[int, ...]]


def _tuple(conv_hyperparameter: IntOrTuple, N: int) -> Tuple[int, ...]:
    """Normalise a conv hyper-parameter (int or N-tuple of int) to an N-tuple.

    Mirrors ``einconv/utils.py:12-48`` including its ``ValueError`` conditions.
    """
    h = conv_hyperparameter
    if isinstance(h, int):
        return (h,) * N
    if not isinstance(h, tuple):
        raise ValueError(
            "Convolution hyperparameters must be specified as int or tuple of int."
            f" Got {h}."
        )
    if len(h) != N:
        raise ValueError(
            f"Convolution hyperparameters must be specified as tuple of length {N}."
            f" Got tuple of length {len(h)}."
        )
    if not all(isinstance(entry, int) for entry in h):
        raise ValueError(f"All hyperparameters must be integers. Got {h}.")
    return h


def get_conv_paddings(
    kernel_size: int, stride: int, padding: Union[int, str], dilation: int
) -> Tuple[int, int]:
    """Left/right zero padding along one dimension (``einconv/utils.py:97-132``).

    ``'same'`` splits ``dilation * (kernel_size - 1)`` with the smaller half on
    the left and is only defined for unit stride.
    """
    if not isinstance(padding, str):
        return padding, padding
    if padding == "valid":
        return 0, 0
    if padding == "same":
        if stride != 1:
            raise ValueError(
                "padding='same' is not supported for strided convolutions."
            )
        total = dilation * (kernel_size - 1)
        left = total // 2
        return left, total - left
    raise ValueError(f"Unknown string-value for padding: {padding!r}.")


def get_conv_output_size(
    input_size: int,
    kernel_size: int,
    stride: int,
    padding: Union[int, str],
    dilation: int,
) -> int:
    """Number of output positions along one dimension (``einconv/utils.py:135-160``)."""
    left, right = get_conv_paddings(kernel_size, stride, padding, dilation)
    span = kernel_size + (kernel_size - 1) * (dilation - 1)
    # integer floor division == floor of the true quotient for stride > 0
    return 1 + (input_size + left + right - span) // stride


def get_conv_input_size(
    output_size: int,
    kernel_size: int,
    stride: int,
    padding: int,
    output_padding: int,
    dilation: int,
) -> int:
    """Input size that maps to ``output_size`` (``einconv/utils.py:163-205``)."""
    input_size = (
        (output_size - 1) * stride
        - 2 * padding
        + dilation * (kernel_size - 1)
        + 1
        + output_padding
    )
    recomputed = get_conv_output_size(input_size, kernel_size, stride, padding, dilation)
    if recomputed != output_size:
        raise ValueError(
            f"Output size {output_size} does not match re-computed output size "
            f"{recomputed}."
        )
    return input_size


def compare_attributes(obj1: Any, obj2: Any, attributes: List[str]) -> None:
    """Raise ``ValueError`` if any of the named attributes differ."""
    for attr in attributes:
        a1, a2 = getattr(obj1, attr), getattr(obj2, attr)
        if a1 != a2:
            raise ValueError(f"{attr!r} attribute does not match: {a1} ≠ {a2}")


def sync_parameters(module1: Module, module2: Module) -> None:
    """Copy ``module1``'s parameter values into ``module2`` (``einconv/utils.py:51-76``)."""
    names1 = {n for n, _ in module1.named_parameters()}
    names2 = {n for n, _ in module2.named_parameters()}
    if names1 != names2:
        raise ValueError(
            f"Module parameters have different names: {names1} ≠ {names2}"
        )
    for name in names1:
        p1, p2 = getattr(module1, name), getattr(module2, name)
        if p1 is not None or p2 is not None:
            compare_attributes(p1, p2, ["shape", "dtype", "device"])
            p2.data = p1.data.clone()


def resolve_geometry(
    input_size: Sequence[int],
    kernel_size: IntOrTuple,
    stride: IntOrTuple = 1,
    padding: Union[int, str, Tuple[int, ...]] = 0,
    dilation: IntOrTuple = 1,
):
    """Per-dimension ``(I, K, S, D, P_left, O)`` tuples for an N-d convolution.

    One string padding applies to every dimension
    (``einconv/expressions/utils.py:47,55``).
    """
    N = len(input_size)
    t_in = tuple(int(i) for i in input_size)
    t_k = _tuple(kernel_size, N)
    t_s = _tuple(stride, N)
    t_d = _tuple(dilation, N)
    t_p = padding if isinstance(padding, str) else _tuple(padding, N)
    p_left, out = [], []
    for d in range(N):
        p = t_p if isinstance(t_p, str) else t_p[d]
        p_left.append(get_conv_paddings(t_k[d], t_s[d], p, t_d[d])[0])
        out.append(get_conv_output_size(t_in[d], t_k[d], t_s[d], p, t_d[d]))
    return t_in, t_k, t_s, t_d, tuple(p_left), tuple(out)
