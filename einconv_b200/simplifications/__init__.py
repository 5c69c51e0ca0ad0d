"""Simplification of einsum expressions of convolution tensor networks."""

from typing import List, Tuple, Union

from torch import Tensor
from torch.nn import Parameter

from einconv_b200.simplifications.opt import Identity, TensorNetwork

__all__ = ["simplify", "TensorNetwork", "Identity"]


def simplify(
    equation: str, operands: List[Union[Tensor, Parameter]]
) -> Tuple[str, List[Union[Tensor, Parameter]]]:
    """Simplify an einsum expression (``einconv/simplifications/__init__.py:11-26``).

    Returns the rewritten equation and operands; the contraction result is unchanged
    up to a reshape.
    """
    network = TensorNetwork(equation, list(operands))
    network.simplify()
    return network.generate_expression()
