"""Autograd parity: the reference evaluates every network with ``torch.einsum`` and is therefore
differentiable to any order (``einconv/modules/conv.py:173-182``, ``README.md:64-73``).  The native
kernels get their derivatives from ``torch.autograd.Function``s whose backward passes are again
native, differentiable calls (``einconv_b200/evaluator.py``).  Checked in fp64 with finite differences
(``gradcheck`` / ``gradgradcheck``) and against ``torch.nn.functional``'s own autograd.
"""

import pytest
import torch
import torch.nn.functional as F
from torch.autograd import gradcheck, gradgradcheck

from einconv_b200 import evaluator
from einconv_b200.expressions import (
    conv_transposeNd_kfac_reduce,
    convNd_input_vjp,
    convNd_kfac_reduce,
    convNd_kfc,
    convNd_weight_vjp,
)
from einconv_b200.functionals import convNd, unfoldNd, unfoldNd_transpose

pytestmark = pytest.mark.gpu
DEV = "cuda"
KW = dict(stride=2, padding=1, dilation=1, groups=2)


def _rand(*shape, grad=True):
    return torch.rand(*shape, dtype=torch.float64, device=DEV, requires_grad=grad)


def test_convnd_first_and_second_order_gradients_fp64():
    torch.manual_seed(0)
    x, w, b = _rand(2, 4, 7, 6), _rand(6, 2, 3, 3), _rand(6)
    fn = lambda x, w, b: convNd(x, w, b, **KW)  # noqa: E731
    assert gradcheck(fn, (x, w, b), eps=1e-6, atol=1e-6)
    assert gradgradcheck(fn, (x, w, b), eps=1e-6, atol=1e-5)


def test_double_backward_matches_torch_conv2d():
    """Gradient penalty style: d/d(x,w) of |dL/dx|^2 + |dL/dw|^2 (needs the backward's own graph)."""
    torch.manual_seed(1)
    x, w = _rand(2, 4, 8, 8), _rand(6, 2, 3, 3)
    v = torch.rand(2, 6, 4, 4, dtype=torch.float64, device=DEV)

    def penalty(conv):
        y = conv(x, w)
        gx, gw = torch.autograd.grad(y, (x, w), v, create_graph=True)
        return torch.autograd.grad((gx**2).sum() + (gw**2).sum(), (x, w))

    ours = penalty(lambda x, w: convNd(x, w, None, **KW))
    ref = penalty(lambda x, w: F.conv2d(x, w, None, **KW))
    for a, b in zip(ours, ref):
        torch.testing.assert_close(a, b, rtol=1e-9, atol=1e-9)


def test_vjp_expressions_are_differentiable_through_torch_einsum():
    torch.manual_seed(2)
    x, w = _rand(2, 4, 7, 6), _rand(6, 2, 3, 3)
    v = _rand(2, 6, 4, 3)

    def input_vjp(w, v):
        eq, ops, shape = convNd_input_vjp.einsum_expression(w, v, (7, 6), **KW)
        return torch.einsum(eq, *ops).reshape(shape)

    def weight_vjp(x, v):
        eq, ops, shape = convNd_weight_vjp.einsum_expression(x, v, 3, **KW)
        return torch.einsum(eq, *ops).reshape(shape)

    assert input_vjp(w, v).grad_fn is not None and weight_vjp(x, v).grad_fn is not None
    assert gradcheck(input_vjp, (w, v), eps=1e-6, atol=1e-6)
    assert gradcheck(weight_vjp, (x, v), eps=1e-6, atol=1e-6)
    assert gradgradcheck(weight_vjp, (x, v), eps=1e-6, atol=1e-5)


@pytest.mark.parametrize("kind", ["kfc", "kfac_reduce"])
def test_factor_gradients(kind):
    """A factor of an input that requires grad keeps its graph: gradient against the definition
    through ``F.unfold``."""
    torch.manual_seed(3)
    x = _rand(3, 4, 6, 5)
    kw = dict(stride=1, padding=1, dilation=1, groups=2)
    mod = convNd_kfc if kind == "kfc" else convNd_kfac_reduce

    def ours(x):
        eq, ops, shape = mod.einsum_expression(x, 3, **kw)
        return torch.einsum(eq, *ops).reshape(shape)

    # the reference's scale is ``Tensor([1 / B]).to(x.dtype)``: rounded through fp32 (convNd_kfc.py:105)
    scale = float(torch.tensor([1.0 / 3]))

    def definition(x):
        u = F.unfold(x, 3, padding=1).reshape(3, 2, 18, -1)
        if kind == "kfc":
            return torch.einsum("ngik,ngjk->gij", u, u) * scale
        m = u.sum(-1)
        return torch.einsum("ngi,ngj->gij", m, m) * float(torch.tensor([1.0 / (3 * 30**2)]))

    out = ours(x)
    assert out.grad_fn is not None
    torch.testing.assert_close(out, definition(x), rtol=1e-10, atol=1e-12)
    c = torch.rand_like(out)
    (g1,) = torch.autograd.grad(out, x, c)
    (g2,) = torch.autograd.grad(definition(x), x, c)
    torch.testing.assert_close(g1, g2, rtol=1e-9, atol=1e-12)
    assert gradcheck(ours, (x,), eps=1e-6, atol=1e-6)
    assert gradgradcheck(ours, (x,), eps=1e-6, atol=1e-5)
    with pytest.raises(RuntimeError):
        (evaluator.kfc if kind == "kfc" else evaluator.kfac_reduce)(x, 3, padding=1, groups=2, out=torch.empty_like(out))


def test_unfold_and_fold_are_each_others_derivative():
    torch.manual_seed(4)
    x = _rand(2, 3, 7, 6)
    fn = lambda x: unfoldNd(x, 3, dilation=1, padding=1, stride=2)  # noqa: E731
    assert gradcheck(fn, (x,), eps=1e-6, atol=1e-6)
    assert gradgradcheck(fn, (x,), eps=1e-6, atol=1e-6)


def test_transpose_networks_are_differentiable():
    torch.manual_seed(5)
    x = _rand(2, 4, 4, 3)
    fn = lambda x: unfoldNd_transpose(x, 3, stride=2, padding=1, output_padding=1)  # noqa: E731
    assert fn(x).grad_fn is not None
    assert gradcheck(fn, (x,), eps=1e-6, atol=1e-6)

    def factor(x):
        eq, ops, shape = conv_transposeNd_kfac_reduce.einsum_expression(x, 3, stride=2, padding=1, output_padding=1,
                                                                        groups=2)
        return torch.einsum(eq, *ops).reshape(shape)

    ref = evaluator.kfac_reduce_transpose(x.detach(), 3, stride=2, padding=1, output_padding=1, groups=2)
    torch.testing.assert_close(factor(x), ref, rtol=1e-10, atol=1e-14)
    assert gradcheck(factor, (x,), eps=1e-6, atol=1e-6)


def test_fp32_training_step_uses_native_backward_kernels():
    """No graph is recorded when nothing requires grad; with grad, backward launches native kernels."""
    from einconv_b200 import _native as nat

    x = torch.rand(2, 8, 10, 10, device=DEV)
    w = torch.rand(8, 8, 3, 3, device=DEV, requires_grad=True)
    assert convNd(x, w.detach(), padding=1).grad_fn is None
    y = convNd(x, w, padding=1)
    before = nat.launch_count()
    y.sum().backward()
    assert nat.launch_count() > before and w.grad is not None


# =============================================================================================
# user-composed networks with index patterns
# =============================================================================================
def _dense_einsum(eq, ops):
    return torch.einsum(eq, *evaluator.plain(ops))


def test_user_composed_networks_apply_patterns_as_gathers(monkeypatch):
    """Networks built by hand from ``einconv_b200.index_pattern`` never materialise the one-hot tensor when the
    pattern's input (or output) index is contracted with one data operand: same values as the dense route."""
    import einconv_b200
    from einconv_b200 import conv_index_pattern as cip

    torch.manual_seed(6)
    x = torch.rand(3, 5, 11, 9, dtype=torch.float64, device=DEV)
    w = torch.rand(4, 5, 3, 2, dtype=torch.float64, device=DEV)
    pat_h = einconv_b200.index_pattern(11, 3, stride=2, padding=1, dilation=1, device=DEV, dtype=x.dtype)
    pat_w = einconv_b200.index_pattern(9, 2, stride=1, padding=0, dilation=2, device=DEV, dtype=x.dtype)
    calls = []
    real = cip.IndexPattern.materialize
    monkeypatch.setattr(cip.IndexPattern, "materialize", lambda self: calls.append(1) or real(self))

    # a convolution written by hand, with an extra per-channel gate the library has no plan for
    gate = torch.rand(5, dtype=torch.float64, device=DEV)
    eq = "ncab,kxa,lyb,dckl,c->ndxy"
    got = torch.einsum(eq, x, pat_h, pat_w, w, gate)
    assert not calls, "the index patterns were written out"
    want = F.conv2d(x * gate.view(1, -1, 1, 1), w, stride=(2, 1), padding=(1, 0), dilation=(1, 2))
    torch.testing.assert_close(got, want, rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(got, _dense_einsum(eq, [x, pat_h, pat_w, w, gate]), rtol=1e-10, atol=1e-12)

    # second-moment of the unfolded input along one axis only (a "lens" onto the network)
    calls.clear()
    eq2 = "ncab,kxa,lxa_->nckl".replace("a_", "e").replace("ncab", "ncab")
    eq2 = "ncab,kxa,lxe,nceb->nckl"
    got2 = torch.einsum(eq2, x, pat_h, pat_h, x)
    assert not calls
    torch.testing.assert_close(got2, _dense_einsum(eq2, [x, pat_h, pat_h, x]), rtol=1e-10, atol=1e-12)

    # transpose direction: the OUTPUT index is contracted (a transpose convolution written by hand)
    calls.clear()
    y = torch.rand(3, 4, 6, dtype=torch.float64, device=DEV)   # [n, d, x]: 6 = output size of pat_h
    wt = torch.rand(4, 5, 3, dtype=torch.float64, device=DEV)   # [d, c, k]
    eq3 = "ndx,kxa,dck->nca"
    got3 = torch.einsum(eq3, y, pat_h, wt)
    assert not calls
    want3 = F.conv_transpose1d(y, wt, stride=2, padding=1)
    torch.testing.assert_close(got3, want3[..., :11], rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(got3, _dense_einsum(eq3, [y, pat_h, wt]), rtol=1e-10, atol=1e-12)

    # a pattern whose indices all survive is written out (nothing to apply it to): still correct
    calls.clear()
    got4 = torch.einsum("kxa,a->kx", pat_h, torch.ones(11, dtype=torch.float64, device=DEV))
    torch.testing.assert_close(got4, _dense_einsum("kxa,a->kx", [pat_h, torch.ones(11, dtype=torch.float64, device=DEV)]))

    # gradients flow through the gathers
    xg = x.clone().requires_grad_()
    out = torch.einsum(eq, xg, pat_h, pat_w, w, gate)
    (gx,) = torch.autograd.grad(out.sum(), xg)
    xr = x.clone().requires_grad_()
    (gr,) = torch.autograd.grad(F.conv2d(xr * gate.view(1, -1, 1, 1), w, stride=(2, 1), padding=(1, 0),
                                         dilation=(1, 2)).sum(), xr)
    torch.testing.assert_close(gx, gr, rtol=1e-10, atol=1e-12)
