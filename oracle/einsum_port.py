"""ORACLE (test infrastructure): the reference's CPU algorithm, restated.

f-dangel/einconv evaluates every operation as ONE ``torch.einsum`` over the data tensors and one
dense one-hot *index pattern* per spatial dimension (``einconv/functionals/conv.py:61``,
``einconv/functionals/unfold.py:56``, ``README.md:64-73``).  The arithmetic therefore lives in a
third-party dependency that is not part of /root/reference: ``torch`` (unpinned in the reference's
``setup.cfg:33-35``; this image: torch 2.11.0, MKL, no opt_einsum => left-to-right pairwise
``bmm``).  This module rebuilds exactly that computation:

* ``pattern``            closed-form index pattern, ``conv_index_pattern.py:120-134``
* ``*_expression``       the un-simplified networks of ``expressions/convNd_*.py`` (equation strings
                         as the reference builds them before ``translate_to_torch``)
* ``evaluate``           ``torch.einsum(equation, *operands).reshape(shape)``

The reference additionally rewrites the network (``simplify=True``) before the einsum; that only
regroups / squeezes indices (and removes the patterns of dense / 1x1 convolutions), it does not
change the contraction algorithm, so the un-simplified network is used here and the difference in
CPU time was measured once with the real reference ("CPU baseline").

``bench.py`` times this module as the CPU baseline (``cpu_baseline.kind = "port"``).
"""

from typing import List, Sequence, Tuple, Union

import torch
from torch import Tensor

from oracle.direct import output_size, paddings


def _tup(h, n):
    return (h,) * n if isinstance(h, int) else tuple(h)


def pattern(I: int, K: int, S: int, P: Union[int, str], D: int, dtype, device="cpu") -> Tensor:
    """Pi[k, o, i] = 1 iff i == S*o + D*k - P_left and 0 <= i < I."""
    O = output_size(I, K, S, P, D)
    p_left, _ = paddings(K, S, P, D)
    pi = torch.zeros(K, O, I, dtype=torch.bool, device=device)
    o = torch.arange(O, device=device)
    for k in range(K):
        i = -p_left + k * D + S * o
        ok = (i >= 0) & (i < I)
        pi[k, o[ok], i[ok]] = True
    return pi.to(dtype)


def patterns(input_size: Sequence[int], kernel_size, stride, padding, dilation, dtype, device="cpu"):
    n = len(input_size)
    ks, ss, ds = _tup(kernel_size, n), _tup(stride, n), _tup(dilation, n)
    ps = (padding,) * n if isinstance(padding, (int, str)) else tuple(padding)
    return [pattern(input_size[d], ks[d], ss[d], ps[d], ds[d], dtype, device) for d in range(n)]


_LETTERS = "abcdefghijklmnopqrstuvwxyz"


def _compile(lhs: List[List[str]], rhs: List[str]) -> str:
    names = []
    for term in lhs:
        for name in term:
            if name not in names:
                names.append(name)
    if len(names) > len(_LETTERS):
        raise ValueError("Ran out of letters.")
    to = {name: _LETTERS[i] for i, name in enumerate(names)}
    return ",".join("".join(to[n] for n in term) for term in lhs) + "->" + "".join(to[n] for n in rhs)


def _ix(prefix, n, suffix=""):
    return [f"{prefix}{d}{suffix}" for d in range(n)]


def _pat_terms(n, ks="", os_="", is_=""):
    return [[f"k{d}{ks}", f"o{d}{os_}", f"i{d}{is_}"] for d in range(n)]


def forward_expression(x, w, stride=1, padding=0, dilation=1, groups=1):
    n = x.dim() - 2
    pats = patterns(x.shape[2:], tuple(w.shape[2:]), stride, padding, dilation, w.dtype, w.device)
    eq = _compile(
        [["n", "g", "ci", *_ix("i", n)], *_pat_terms(n), ["g", "co", "ci", *_ix("k", n)]],
        ["n", "g", "co", *_ix("o", n)],
    )
    xg = x.reshape(x.shape[0], groups, x.shape[1] // groups, *x.shape[2:])
    wg = w.reshape(groups, w.shape[0] // groups, *w.shape[1:])
    shape = (x.shape[0], w.shape[0], *[p.shape[1] for p in pats])
    return eq, [xg, *pats, wg], shape


def input_vjp_expression(w, v, input_size, stride=1, padding=0, dilation=1, groups=1):
    n = w.dim() - 2
    input_size = _tup(input_size, n)
    pats = patterns(input_size, tuple(w.shape[2:]), stride, padding, dilation, w.dtype, w.device)
    eq = _compile(
        [["n", "g", "co", *_ix("o", n)], *_pat_terms(n), ["g", "co", "ci", *_ix("k", n)]],
        ["n", "g", "ci", *_ix("i", n)],
    )
    vg = v.reshape(v.shape[0], groups, v.shape[1] // groups, *v.shape[2:])
    wg = w.reshape(groups, w.shape[0] // groups, *w.shape[1:])
    shape = (v.shape[0], groups * w.shape[1], *input_size)
    return eq, [vg, *pats, wg], shape


def weight_vjp_expression(x, v, kernel_size, dilation=1, padding=0, stride=1, groups=1):
    n = x.dim() - 2
    ks = _tup(kernel_size, n)
    pats = patterns(x.shape[2:], ks, stride, padding, dilation, x.dtype, x.device)
    eq = _compile(
        [["n", "g", "ci", *_ix("i", n)], *_pat_terms(n), ["n", "g", "co", *_ix("o", n)]],
        ["g", "co", "ci", *_ix("k", n)],
    )
    xg = x.reshape(x.shape[0], groups, x.shape[1] // groups, *x.shape[2:])
    vg = v.reshape(v.shape[0], groups, v.shape[1] // groups, *v.shape[2:])
    shape = (v.shape[1], x.shape[1] // groups, *ks)
    return eq, [xg, *pats, vg], shape


def unfold_expression(x, kernel_size, stride=1, padding=0, dilation=1):
    n = x.dim() - 2
    ks = _tup(kernel_size, n)
    pats = patterns(x.shape[2:], ks, stride, padding, dilation, x.dtype, x.device)
    eq = _compile(
        [["n", "ci", *_ix("i", n)], *_pat_terms(n)],
        ["n", "ci", *_ix("k", n), *_ix("o", n)],
    )
    kt = 1
    for k in ks:
        kt *= k
    ot = 1
    for p in pats:
        ot *= p.shape[1]
    return eq, [x, *pats], (x.shape[0], x.shape[1] * kt, ot)


def _factor_expression(x, kernel_size, stride, padding, dilation, groups, shared_o: bool):
    n = x.dim() - 2
    ks = _tup(kernel_size, n)
    pats = patterns(x.shape[2:], ks, stride, padding, dilation, x.dtype, x.device)
    eq = _compile(
        [
            ["n", "g", "ci", *_ix("i", n)],
            *_pat_terms(n),
            *_pat_terms(n, ks="_", os_="" if shared_o else "_", is_="_"),
            ["n", "g", "ci_", *_ix("i", n, "_")],
            ["s"],
        ],
        ["g", "ci", *_ix("k", n), "ci_", *_ix("k", n, "_")],
    )
    xg = x.reshape(x.shape[0], groups, x.shape[1] // groups, *x.shape[2:])
    kt = 1
    for k in ks:
        kt *= k
    ot = 1
    for p in pats:
        ot *= p.shape[1]
    B = x.shape[0]
    scale = 1.0 / B if shared_o else 1.0 / (B * ot**2)
    s = Tensor([scale]).to(x.device).to(x.dtype)
    A = x.shape[1] // groups * kt
    return eq, [xg, *pats, *pats, xg, s], (groups, A, A)


def kfc_expression(x, kernel_size, stride=1, padding=0, dilation=1, groups=1):
    return _factor_expression(x, kernel_size, stride, padding, dilation, groups, shared_o=True)


def kfac_reduce_expression(x, kernel_size, stride=1, padding=0, dilation=1, groups=1):
    return _factor_expression(x, kernel_size, stride, padding, dilation, groups, shared_o=False)


def evaluate(expression: Tuple[str, List[Tensor], Tuple[int, ...]]) -> Tensor:
    """The reference's hot call: ``torch.einsum(equation, *operands).reshape(shape)``."""
    equation, operands, shape = expression
    return torch.einsum(equation, *operands).reshape(shape)


def convNd(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1) -> Tensor:
    """einconv/functionals/conv.py:52-69."""
    out = evaluate(forward_expression(x, w, stride, padding, dilation, groups))
    if bias is not None:
        out += bias.reshape(1, -1, *([1] * (x.dim() - 2))).expand_as(out)
    return out


def unfoldNd(x, kernel_size, dilation=1, padding=0, stride=1) -> Tensor:
    """einconv/functionals/unfold.py:48-56."""
    return evaluate(unfold_expression(x, kernel_size, stride, padding, dilation))
