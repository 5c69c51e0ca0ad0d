"""Robustness of the dispatch: tensors whose ``data_ptr`` is not 16-byte aligned.

A contiguous batch-shard view such as ``x_all[1:]`` of a ``[B, 3, 7, 7]`` tensor starts 588 bytes into
its storage.  The tensor-core routes move data with TMA boxes and 16-byte vector accesses; the
dispatcher must pick a route that accepts the pointer instead of failing (``csrc/api.cu``).
"""

import numpy as np
import pytest
import torch

from einconv_b200 import evaluator
from einconv_b200.functionals import convNd
from oracle import direct

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(params=["fp32", "tf32"])
def math_mode(request):
    previous = evaluator.get_fp32_math()
    evaluator.set_fp32_math(request.param)
    yield request.param
    evaluator.set_fp32_math(previous)


def _tol(mode):
    return dict(rtol=2e-3, atol=2e-3) if mode == "tf32" else dict(rtol=5e-5, atol=1e-5)


@pytest.mark.parametrize("shape,k,pad", [((5, 3, 7, 7), 3, 1), ((3, 33, 9, 9), 3, 1), ((3, 65, 5, 5), 1, 0)])
def test_factors_of_a_misaligned_batch_shard(shape, k, pad, math_mode):
    torch.manual_seed(0)
    x_all = torch.rand(*shape, device=DEV)
    x = x_all[1:]
    assert x.is_contiguous() and x.data_ptr() % 16 != 0
    want_k = direct.kfc(x.cpu().numpy(), k, padding=pad)
    want_r = direct.kfac_reduce(x.cpu().numpy(), k, padding=pad)
    np.testing.assert_allclose(evaluator.kfc(x, k, padding=pad).cpu().numpy(), want_k, **_tol(math_mode))
    np.testing.assert_allclose(evaluator.kfac_reduce(x, k, padding=pad).cpu().numpy(), want_r, **_tol(math_mode))


def test_convolution_networks_with_misaligned_operands(math_mode):
    torch.manual_seed(1)
    x_all = torch.rand(3, 33, 9, 9, device=DEV)
    w_all = torch.rand(1 + 64 * 33 * 9, device=DEV)
    x = x_all[1:]
    w = w_all[1:].view(64, 33, 3, 3)
    assert x.data_ptr() % 16 != 0 and w.data_ptr() % 16 != 0
    kw = dict(stride=1, padding=1, dilation=1, groups=1)
    xn, wn = x.cpu().numpy(), w.cpu().numpy()
    y = convNd(x, w, None, **kw)
    np.testing.assert_allclose(y.cpu().numpy(), direct.conv_forward(xn, wn, None, **kw), **_tol(math_mode))
    v = torch.rand(1 + y.numel(), device=DEV)[1:].view_as(y)
    gx = evaluator.conv_input_vjp_raw(w, v, (9, 9), 1, 1, 1, 1)
    np.testing.assert_allclose(gx.cpu().numpy(), direct.conv_input_vjp(wn, v.cpu().numpy(), (9, 9), **kw),
                               **_tol(math_mode))
    gw = evaluator.conv_weight_vjp_raw(x, v, 3, 1, 1, 1, 1)
    np.testing.assert_allclose(gw.cpu().numpy(), direct.conv_weight_vjp(xn, v.cpu().numpy(), 3, dilation=1, padding=1,
                                                                         stride=1, groups=1), **_tol(math_mode))


def test_misaligned_bf16_operands():
    torch.manual_seed(2)
    x = torch.rand(1 + 2 * 64 * 10 * 10, device=DEV).bfloat16()[1:].view(2, 64, 10, 10)
    w = torch.rand(64, 64, 3, 3, device=DEV).bfloat16()
    assert x.data_ptr() % 16 != 0
    y = convNd(x, w, None, padding=1)
    want = direct.conv_forward(x.double().cpu().numpy(), w.double().cpu().numpy(), None, stride=1, padding=1,
                               dilation=1, groups=1)
    assert np.abs(y.double().cpu().numpy() - want).max() <= 2.0**-7 * np.abs(want).max()
    k = evaluator.kfc(x, 3, padding=1)
    want_k = direct.kfc(x.double().cpu().numpy(), 3, padding=1)
    assert np.abs(k.double().cpu().numpy() - want_k).max() <= 2.0**-7 * np.abs(want_k).max()
