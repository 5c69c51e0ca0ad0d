"""Helpers for building einsum expressions of convolution tensor networks.

Behavioural contract: ``einconv/expressions/utils.py`` — ``create_conv_index_patterns``
(``:12-61``), ``translate_to_torch`` (``:64-111``, exact-string known answers in the
reference's ``test/expressions/test_utils.py:19-25``) and ``get_letters``
(``:114-145``, 26-letter budget).
"""

from typing import Iterable, List, Optional, Tuple, Union

import torch
from torch import Tensor

from einconv_b200.conv_index_pattern import index_pattern
from einconv_b200.utils import _tuple, cpu

_ALPHABET = "abcdefghijklmnopqrstuvwxyz"


def create_conv_index_patterns(
    N: int,
    input_size: Union[int, Tuple[int, ...]],
    kernel_size: Union[int, Tuple[int, ...]],
    stride: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, str, Tuple[int, ...]] = 0,
    dilation: Union[int, Tuple[int, ...]] = 1,
    device: torch.device = cpu,
    dtype: torch.dtype = torch.bool,
) -> List[Tensor]:
    """One (lazy) index pattern per spatial dimension of an ``N``-d convolution.

    A string ``padding`` is shared by all dimensions; everything else may be an int
    or an ``N``-tuple.
    """
    sizes = _tuple(tuple(input_size) if not isinstance(input_size, int) else input_size, N)
    kernels = _tuple(kernel_size, N)
    strides = _tuple(stride, N)
    dilations = _tuple(dilation, N)
    paddings = N * (padding,) if isinstance(padding, str) else _tuple(padding, N)

    return [
        index_pattern(
            sizes[d],
            kernels[d],
            stride=strides[d],
            padding=paddings[d],
            dilation=dilations[d],
            device=device,
            dtype=dtype,
        )
        for d in range(N)
    ]


def translate_to_torch(einops_equation: str) -> str:
    """Turn an einops-style equation (``'row j, j col -> row col'``) into torch syntax.

    Multi-character or non-lowercase index names are replaced by free single letters,
    assigned in order of first appearance on the left-hand side.
    """
    lhs = einops_equation.split("->")[0]
    names: List[str] = []
    for operand in lhs.strip().split(","):
        for name in operand.strip().split(" "):
            if name not in names:
                names.append(name)
    if "" in names:
        raise ValueError(f"Index parsing failed. Equation {einops_equation}: {names}")

    long_names = [n for n in names if len(n) > 1 or not n.islower()]
    short_names = {n for n in names if n not in long_names}
    letter_of = dict(zip(long_names, get_letters(len(long_names), blocked=short_names)))

    # Substitute a name only once no other pending name contains it as a sub-string
    # ('row' must wait for 'row2').
    equation = einops_equation
    pending = list(long_names)
    while pending:
        name = pending.pop(0)
        if any(name in other for other in pending):
            pending.append(name)
            continue
        equation = equation.replace(name, letter_of[name])

    return equation.replace(" ", "")


def get_letters(num_letters: int, blocked: Optional[Iterable[str]] = None) -> List[str]:
    """First ``num_letters`` lowercase letters that are not in ``blocked``."""
    if num_letters == 0:
        return []
    taken = set(blocked) if blocked is not None else set()
    free = [c for c in _ALPHABET if c not in taken]
    if len(free) < num_letters:
        raise ValueError(
            f"Ran out of letters. PyTorch's einsum supports {len(_ALPHABET)} letters."
            f" Requested {num_letters}, blocked: {len(taken)}.)"
        )
    return free[:num_letters]
