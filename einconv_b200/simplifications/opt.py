"""Rewrites of convolution tensor networks given as einsum equation + operands.

Behavioural contract: ``einconv/simplifications/opt.py`` — ``Identity`` (``:16-49``)
and ``TensorNetwork`` (``:52-473``) with the pass order of ``simplify`` (``:96-102``):
down-sampling (``:196-248``) -> dense (``:250-300``) -> identity pruning
(``:190-194, 390-462``) -> parallel-index merging (``:104-146``) -> squeezing
(``:148-171``).  The letters handed out by ``ungroup`` and the order in which rules
fire are observable through the returned equation (the reference asserts exact
strings, ``test/simplifications/test_opt.py:20,35,50,80,189,211,232``), so both are
kept.

In this build the rewritten expression is *not* what runs on the GPU: the native
evaluator works from the convolution geometry and only uses the per-dimension
predicates in :mod:`einconv_b200.simplifications.rules` (the same ones that drive the
rewrites here) to pick kernel specialisations.  The rewrite is kept because
``einconv.simplify`` is public API and ``einsum_expression`` must keep returning a
valid, simplified triple.
"""

from __future__ import annotations

from math import prod
from typing import Dict, List, Sequence, Tuple, Union

import torch
from torch import Tensor
from torch.nn import Parameter

from einconv_b200.conv_index_pattern import index_pattern
from einconv_b200.expressions.utils import get_letters
from einconv_b200.simplifications.rules import is_dense, is_downsampling
from einconv_b200.utils import cpu


class Identity:
    """Placeholder for a Kronecker delta ``delta[i, j]`` of a given size."""

    def __init__(
        self, dim: int, device: torch.device = cpu, dtype: torch.dtype = torch.float32
    ) -> None:
        self._dim = dim
        self._device = device
        self._dtype = dtype

    def to_tensor(self) -> Tensor:
        """Dense identity matrix with the stored size, device and dtype."""
        return torch.eye(self._dim, device=self._device, dtype=self._dtype)

    def __repr__(self) -> str:  # noqa: D105
        return f"Identity({self._dim}, device={self._device}, dtype={self._dtype})"


Operand = Union[Tensor, Parameter, Identity]


def _hyperparams(op) -> Union[Dict, None]:
    return getattr(op, "_pattern_hyperparams", None)


def _shape_standin(op: Tensor) -> Tensor:
    """Storage-free stand-in for a pattern that is about to be replaced.

    Index surgery (reshape / narrow / flatten) on it costs nothing and, for a lazy
    :class:`IndexPattern`, avoids writing out the dense one-hot tensor.
    """
    return torch.empty(tuple(op.shape), dtype=op.dtype, device="meta")


class TensorNetwork:
    """Einsum equation + operands with convolution-aware simplification passes.

    Usage: construct, call :meth:`simplify`, then :meth:`generate_expression`.  The
    contraction result may come out with a different shape than the original one but
    holds the same numbers up to a reshape.
    """

    def __init__(self, equation: str, operands: List[Operand]) -> None:
        lhs, rhs = equation.split("->")
        self.input_indices: List[str] = lhs.split(",")
        self.output_indices: str = rhs
        self.operands: List[Operand] = operands

        self.dims: Dict[str, int] = {}
        for letters, op in zip(self.input_indices, self.operands):
            if isinstance(op, Identity):
                self.dims.update({letter: op._dim for letter in letters})
            elif isinstance(op, Tensor):
                self.dims.update(dict(zip(letters, op.shape)))

        self.output_shape = [self.dims[letter] for letter in self.output_indices]

    # ------------------------------------------------------------------ driver
    def simplify(self) -> None:
        """Apply all rewrite passes in the reference's order."""
        self._simplify_downsampling()
        self._simplify_dense()
        self.prune_identities()
        self.merge_parallel()
        self.squeeze()

    def generate_expression(self) -> Tuple[str, List[Union[Tensor, Parameter]]]:
        """Equation and operands of the current network (deltas written out)."""
        equation = ",".join(self.input_indices) + "->" + self.output_indices
        operands = [
            op.to_tensor() if isinstance(op, Identity) else op for op in self.operands
        ]
        return equation, operands

    def __repr__(self) -> str:  # noqa: D105
        return (
            f"TensorNetwork('{','.join(self.input_indices)}->{self.output_indices}',"
            f" {self.operands}, dims={self.dims})"
        )

    # ------------------------------------------------------------ index surgery
    def _all_index_strings(self) -> List[str]:
        return self.input_indices + [self.output_indices]

    def ungroup(self, idx: str, size: Sequence[int]) -> List[str]:
        """Split index ``idx`` into ``len(size)`` indices of the given sizes.

        The first new index keeps the name ``idx``; the others take the first letters
        not used by any operand.  Returns the new names.
        """
        size = list(size)
        assert self.dims[idx] == prod(size)

        for pos, letters in enumerate(self.input_indices):
            if idx in letters:
                at = letters.index(idx)
                shape = [self.dims[letter] for letter in letters]
                self.operands[pos] = self.operands[pos].reshape(
                    shape[:at] + size + shape[at + 1 :]
                )

        fresh = get_letters(len(size) - 1, blocked=set("".join(self.input_indices)))
        new_idxs = [idx] + fresh
        expanded = "".join(new_idxs)

        self.input_indices = [s.replace(idx, expanded) for s in self.input_indices]
        self.output_indices = self.output_indices.replace(idx, expanded)
        self.dims.update(dict(zip(new_idxs, size)))
        return new_idxs

    def narrow(self, idx: str, start: int, length: int) -> None:
        """Keep only ``[start, start + length)`` of index ``idx`` in every operand."""
        assert start >= 0
        for pos, letters in enumerate(self.input_indices):
            if idx in letters:
                self.operands[pos] = self.operands[pos].narrow(
                    letters.index(idx), start, length
                )
        self.dims[idx] = length

    def group(self, idxs: List[str]) -> str:
        """Fuse adjacent indices ``idxs`` (in that order) into one named ``idxs[0]``."""
        run = "".join(idxs)
        fused = idxs[0]
        fused_dim = prod(int(self.dims[letter]) for letter in idxs)

        for letters in self._all_index_strings():
            if run in letters:
                rest = letters.replace(run, "")
                assert not any(letter in rest for letter in idxs)

        for pos, letters in enumerate(self.input_indices):
            if run in letters:
                at = letters.index(run)
                self.operands[pos] = self.operands[pos].flatten(
                    start_dim=at, end_dim=at + len(idxs) - 1
                )
            self.input_indices[pos] = letters.replace(run, fused)
        self.output_indices = self.output_indices.replace(run, fused)

        self.dims[fused] = fused_dim
        for letter in idxs[1:]:
            del self.dims[letter]
        return fused

    # ----------------------------------------------------- convolution-aware rules
    def _first_pattern(self, predicate) -> int:
        """Position of the first pattern operand satisfying ``predicate`` or -1."""
        for pos, op in enumerate(self.operands):
            hp = _hyperparams(op)
            if hp is not None and predicate(pos, hp):
                return pos
        return -1

    def _simplify_downsampling(self) -> None:
        """Stride > kernel: drop the never-read input pixels, leaving a dense pattern."""

        def applies(pos: int, hp: Dict) -> bool:
            if not is_downsampling(**hp):
                return False
            _, _, i_idx = self.input_indices[pos]
            return i_idx not in self.output_indices

        pos = self._first_pattern(applies)
        while pos >= 0:
            op = self.operands[pos]
            hp = _hyperparams(op)
            I, K, S = hp["input_size"], hp["kernel_size"], hp["stride"]
            device, dtype = op.device, op.dtype
            i_idx = self.input_indices[pos][2]

            self.operands[pos] = _shape_standin(op)  # replaced below anyway
            outer, inner = self.ungroup(i_idx, [I // S, S])
            self.narrow(inner, 0, K)
            self.group([outer, inner])

            self.operands[pos] = index_pattern(
                (K * I) // S,
                K,
                stride=K,
                padding=hp["padding"],
                dilation=hp["dilation"],
                device=device,
                dtype=dtype,
            )
            pos = self._first_pattern(applies)

    def _simplify_dense(self) -> None:
        """Kernel == stride, no overlap: the pattern factorises into two deltas."""

        def applies(pos: int, hp: Dict) -> bool:
            return is_dense(**hp)

        pos = self._first_pattern(applies)
        while pos >= 0:
            op = self.operands[pos]
            assert isinstance(op, Tensor)
            hp = _hyperparams(op)
            I, K = hp["input_size"], hp["kernel_size"]
            device, dtype = op.device, op.dtype
            k_idx, o_idx, i_idx = self.input_indices[pos]

            self.operands[pos] = _shape_standin(op)  # replaced below anyway
            i_outer, i_inner = self.ungroup(i_idx, [I // K, K])

            self.operands[pos] = Identity(I // K, device=device, dtype=dtype)
            self.input_indices[pos] = o_idx + i_outer
            self.operands.insert(pos + 1, Identity(K, device=device, dtype=dtype))
            self.input_indices.insert(pos + 1, k_idx + i_inner)

            pos = self._first_pattern(applies)

    # ---------------------------------------------------------------- identities
    def _identity_positions(self) -> List[int]:
        return [p for p, op in enumerate(self.operands) if isinstance(op, Identity)]

    def prune_identities(self) -> None:
        """Remove Kronecker deltas that act as the identity in the contraction."""
        self._remove_duplicate_identities()
        self._replace_self_contracting_identities()
        self._remove_contracting_identities()

    def _remove_duplicate_identities(self) -> None:
        """Drop repeated deltas over the same index pair.

        Faithful to ``opt.py:443-462`` including its indexing: duplicates are located
        by their rank *among the deltas* and that rank is used as operand position.
        """
        letter_sets = [set(self.input_indices[p]) for p in self._identity_positions()]
        seen: List[set] = []
        drop: List[int] = []
        for letters in letter_sets:
            if letters in seen:
                continue
            seen.append(letters)
            ranks = [r for r, other in enumerate(letter_sets) if other == letters]
            drop.extend(ranks[1:])

        for position in sorted(drop):
            del self.operands[position]
            del self.input_indices[position]

    def _replace_self_contracting_identities(self) -> None:
        """A delta connected to nothing else contracts to the scalar ``dim``."""

        def first_isolated() -> int:
            for pos in self._identity_positions():
                i, j = self.input_indices[pos]
                others = [
                    s for p, s in enumerate(self.input_indices) if p != pos
                ] + [self.output_indices]
                if not any(i in s or j in s for s in others):
                    return pos
            return -1

        pos = first_isolated()
        while pos >= 0:
            i, j = self.input_indices[pos]
            delta = self.operands[pos]
            self.input_indices[pos] = i
            self.operands[pos] = Tensor([delta._dim]).to(delta._device)
            self.dims[i] = 1
            del self.dims[j]
            pos = first_isolated()

    def _remove_contracting_identities(self) -> None:
        """A delta touching another operand only renames one of its indices."""

        def first_connected() -> int:
            for pos in self._identity_positions():
                i, j = self.input_indices[pos]
                others = [s for p, s in enumerate(self.input_indices) if p != pos]
                if any(i in s or j in s for s in others):
                    return pos
            return -1

        pos = first_connected()
        while pos >= 0:
            i, j = self.input_indices[pos]
            del self.operands[pos]
            del self.input_indices[pos]
            self.input_indices = [s.replace(i, j) for s in self.input_indices]
            self.output_indices = self.output_indices.replace(i, j)
            del self.dims[i]
            pos = first_connected()

    # ------------------------------------------------------------ generic rewrites
    def merge_parallel(self) -> None:
        """Fuse index pairs that are adjacent (or absent) everywhere they occur."""

        def fusable(pair: str) -> bool:
            for letters in self._all_index_strings():
                if pair in letters:
                    continue
                if pair[0] in letters or pair[1] in letters:
                    return False
            return True

        def next_pair() -> str:
            for letters in self.input_indices:
                for at in range(len(letters) - 1):
                    pair = letters[at : at + 2]
                    if fusable(pair):
                        return pair
            return ""

        pair = next_pair()
        while pair:
            self.group(list(pair))
            pair = next_pair()

    def squeeze(self) -> None:
        """Eliminate indices of size one (from tensors that have other indices)."""
        unit = [letter for letter, size in self.dims.items() if size == 1]

        for pos in range(len(self.input_indices)):
            for letter in unit:
                letters = self.input_indices[pos]
                if (
                    letter in letters
                    and len(letters) > 1
                    and not isinstance(self.operands[pos], Identity)
                ):
                    self.operands[pos] = self.operands[pos].squeeze(letters.index(letter))
                    self.input_indices[pos] = letters.replace(letter, "")

        for letter in unit:
            self.output_indices = self.output_indices.replace(letter, "")

        alive = set("".join(self._all_index_strings()))
        for letter in list(self.dims):
            if letter not in alive:
                del self.dims[letter]
