"""Shared plumbing of the ``einsum_expression`` generators."""

from typing import Any, Dict, List, Tuple

from torch import Tensor

import einconv_b200
from einconv_b200.evaluator import Equation, Plan, tag_for_dispatch
from einconv_b200.expressions.utils import translate_to_torch


def letters(prefix: str, N: int, suffix: str = "") -> str:
    """``'i0 i1 ... i{N-1}'`` (optionally with a suffix on every name)."""
    return " ".join(f"{prefix}{d}{suffix}" for d in range(N))


def pattern_strs(N: int, k_suffix: str = "", o_suffix: str = "", i_suffix: str = "") -> List[str]:
    """Index strings ``'k_d o_d i_d'`` of the per-dimension patterns."""
    return [f"k{d}{k_suffix} o{d}{o_suffix} i{d}{i_suffix}" for d in range(N)]


def finish(
    kind: str,
    lhs: List[str],
    rhs: str,
    operands: List[Tensor],
    shape: Tuple[int, ...],
    tensors: Dict[str, Tensor],
    hyper: Dict[str, Any],
    simplify: bool,
):
    """Translate, optionally simplify, and attach the native plan to the equation."""
    equation = translate_to_torch("->".join([",".join(lhs), rhs]))
    if simplify:
        equation, operands = einconv_b200.simplify(equation, operands)
        operands = tag_for_dispatch(operands)
    plan = Plan(kind=kind, tensors=tensors, hyper=hyper, shape=shape, simplify=simplify,
                operands=list(operands))
    return Equation(equation, plan), operands, shape
