"""CPU: host-side mirror of the reference interface (geometry, patterns, equations, simplifier,
modules, error behaviour).  Known answers come from the reference's own tests
(``test/expressions/test_utils.py:19-25``, ``test/simplifications/test_opt.py``,
``test/test_utils.py`` + ``test/utils_cases.py``) and from golden data produced by the reference.
"""

import json

import numpy as np
import pytest
import torch
from torch import nn

import einconv_b200
from einconv_b200 import index_pattern, index_pattern_logical
from einconv_b200.expressions import (
    conv_transposeNd_kfac_reduce,
    conv_transposeNd_unfold,
    convNd_forward,
    convNd_input_vjp,
    convNd_kfac_reduce,
    convNd_kfc,
    convNd_unfold,
    convNd_weight_vjp,
)
from einconv_b200.expressions.utils import get_letters, translate_to_torch
from einconv_b200.modules import ConvNd, UnfoldNd
from einconv_b200.simplifications import Identity, TensorNetwork
from einconv_b200.utils import (
    _tuple,
    get_conv_input_size,
    get_conv_output_size,
    get_conv_paddings,
)
from tests import cases as C
from tests.conftest import GOLDEN, forward_out_shape, golden


def dense(ops):
    from einconv_b200.evaluator import plain

    return plain(ops)


# ---- geometry ---------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "kw",
    [
        dict(input_size=11, kernel_size=3, stride=2, padding=1, dilation=1),
        dict(input_size=30, kernel_size=4, stride=3, padding=2, dilation=2),
        dict(input_size=20, kernel_size=5, stride=1, padding="same", dilation=2),
        dict(input_size=20, kernel_size=4, stride=1, padding="same", dilation=1),
        dict(input_size=9, kernel_size=3, stride=2, padding="valid", dilation=1),
    ],
)
def test_output_size_matches_conv1d(kw):
    x = torch.zeros(1, 1, kw["input_size"])
    w = torch.zeros(1, 1, kw["kernel_size"])
    truth = nn.functional.conv1d(x, w, stride=kw["stride"], padding=kw["padding"], dilation=kw["dilation"])
    assert get_conv_output_size(**kw) == truth.shape[2]


def test_same_padding_is_asymmetric_for_even_spans():
    assert get_conv_paddings(4, 1, "same", 1) == (1, 2)
    assert get_conv_paddings(5, 1, "same", 2) == (4, 4)
    with pytest.raises(ValueError):
        get_conv_paddings(3, 2, "same", 1)
    with pytest.raises(ValueError):
        get_conv_paddings(3, 1, "full", 1)


@pytest.mark.parametrize("kw", [
    dict(output_size=20, kernel_size=3, stride=2, padding=1, output_padding=0, dilation=1),
    dict(output_size=7, kernel_size=4, stride=3, padding=2, output_padding=2, dilation=2),
])
def test_input_size_matches_conv_transpose1d(kw):
    x = torch.zeros(1, 1, kw["output_size"])
    w = torch.zeros(1, 1, kw["kernel_size"])
    truth = nn.functional.conv_transpose1d(
        x, w, stride=kw["stride"], padding=kw["padding"], output_padding=kw["output_padding"],
        dilation=kw["dilation"],
    )
    assert get_conv_input_size(**kw) == truth.shape[2]


def test_tuple_errors():
    assert _tuple(3, 2) == (3, 3)
    assert _tuple((1, 2), 2) == (1, 2)
    for bad in ([1, 2], (1, 2, 3), (1, 2.0)):
        with pytest.raises(ValueError):
            _tuple(bad, 2)


# ---- equation plumbing (exact known answers of the reference) ----------------------------------
def test_translate_to_torch_known_answers():
    assert translate_to_torch("row j, j col -> row col") == "aj,jb->ab"
    assert translate_to_torch("a b, b c -> a c") == "ab,bc->ac"
    # 'row' is a sub-string of 'row2'
    assert translate_to_torch("row row2, row2 col -> row col") == "ab,bc->ac"
    with pytest.raises(ValueError):
        translate_to_torch("a  b, b c -> a c")


def test_get_letters():
    assert get_letters(0) == []
    assert get_letters(3) == ["a", "b", "c"]
    assert get_letters(2, blocked={"a", "c"}) == ["b", "d"]
    with pytest.raises(ValueError):
        get_letters(27)
    with pytest.raises(ValueError):
        get_letters(3, blocked=set("abcdefghijklmnopqrstuvwx"))


# ---- index pattern ------------------------------------------------------------------------------
@pytest.mark.parametrize("n", range(len(C.INDEX_PATTERN_CASES)))
@pytest.mark.parametrize("dtype", [torch.bool, torch.float32, torch.int32])
def test_index_pattern_matches_reference(n, dtype):
    kw = C.INDEX_PATTERN_CASES[n]
    data = np.load(GOLDEN / "index_pattern.npz")
    shape = tuple(data[f"{n}.shape"])
    truth = np.unpackbits(data[f"{n}.pattern"])[: int(np.prod(shape))].reshape(shape).astype(bool)
    pat = index_pattern(**kw, dtype=dtype)
    assert tuple(pat.shape) == shape and pat.dtype == dtype
    assert pat._pattern_hyperparams["input_size"] == kw["input_size"]
    assert np.array_equal(pat.materialize().numpy().astype(bool), truth)
    logical = index_pattern_logical(**kw, dtype=dtype)
    assert np.array_equal(logical.numpy().astype(bool), truth)


def test_index_pattern_is_lazy_and_behaves_like_a_tensor():
    pat = index_pattern(50, 3, stride=2, padding=1, dtype=torch.float32)
    assert pat._dense is None
    assert pat.shape == (3, 25, 50) and pat.dim() == 3 and pat.device.type == "cpu"
    assert pat._dense is None  # metadata does not write the one-hot tensor
    assert float(pat.sum()) == float(pat.materialize().sum())
    assert (pat.reshape(3, -1).shape) == (3, 25 * 50)
    with pytest.raises(ValueError):
        index_pattern(10, 3, stride=2, padding="same")


# ---- simplifier ----------------------------------------------------------------------------------
def test_prune_identities_known_answers():
    torch.manual_seed(0)
    v = torch.rand(10)
    tn = TensorNetwork("ij,j->i", [Identity(10), v])
    tn.prune_identities()
    assert tn.generate_expression()[0] == "j->j"
    tn = TensorNetwork("ij,jk,k->i", [Identity(10), Identity(10), v])
    tn.prune_identities()
    assert tn.generate_expression()[0] == "k->k"
    tn = TensorNetwork("ij,ij,j->i", [Identity(10), Identity(10), v])
    tn._remove_duplicate_identities()
    assert tn.generate_expression()[0] == "ij,j->i"
    tn = TensorNetwork("ij,kl->ijkl", [Identity(10), Identity(10)])
    tn.prune_identities()
    assert tn.generate_expression()[0] == "ij,kl->ijkl"
    tn = TensorNetwork("ij,kl->ij", [Identity(10), Identity(10)])
    tn._replace_self_contracting_identities()
    eq, ops = tn.generate_expression()
    assert eq == "ij,k->ij" and float(ops[1]) == 10.0


def test_ungroup_group_narrow():
    torch.manual_seed(0)
    A, B = torch.rand(3, 10), torch.rand(10, 4)
    truth = A @ B
    tn = TensorNetwork("ij,jk->ik", [A, B])
    new = tn.ungroup("j", [2, 5])
    eq, ops = tn.generate_expression()
    assert new == ["j", "a"] and eq == "ija,jak->ik"
    assert ops[0].shape == (3, 2, 5) and ops[1].shape == (2, 5, 4)
    assert torch.allclose(torch.einsum(eq, *ops), truth)
    tn.group(["j", "a"])
    assert tn.generate_expression()[0] == "ij,jk->ik"
    tn.narrow("j", 0, 4)
    eq, ops = tn.generate_expression()
    assert torch.allclose(torch.einsum(eq, *ops), A[:, :4] @ B[:4])


def test_simplify_dense_and_downsampling_patterns():
    # mirrors test/simplifications/test_opt.py:112-170
    pattern = index_pattern(input_size=10, kernel_size=2, stride=2)
    truth = pattern.materialize().float()
    tn = TensorNetwork("koi->koi", [pattern])
    tn._simplify_dense()
    assert len(tn.operands) == 2 and all(isinstance(op, Identity) for op in tn.operands)
    eq, ops = tn.generate_expression()
    assert torch.equal(torch.einsum(eq, *[o.float() for o in ops]).reshape_as(truth), truth)

    torch.manual_seed(0)
    v = torch.rand(9)
    pattern = index_pattern(input_size=9, kernel_size=2, stride=3, dtype=torch.float32)
    truth = torch.einsum("i,koi->ko", v, pattern.materialize())
    tn = TensorNetwork("i,koi->ko", [v, pattern])
    tn._simplify_downsampling()
    eq, ops = tn.generate_expression()
    assert len(ops) == 2 and eq == "i,koi->ko"
    assert torch.allclose(torch.einsum(eq, *dense(ops)).reshape_as(truth), truth)
    tn._simplify_dense()
    tn.prune_identities()
    eq, ops = tn.generate_expression()
    assert len(ops) == 1
    assert torch.allclose(torch.einsum(eq, *dense(ops)).reshape_as(truth), truth)


def test_squeeze_and_merge_parallel_known_answers():
    # mirrors test/simplifications/test_opt.py:173-233
    torch.manual_seed(0)
    A, B, v = torch.rand(1, 1), torch.rand(1, 9), torch.rand(9)
    truth = torch.einsum("ij,jk,k->i", A, B, v)
    tn = TensorNetwork("ij,jk,k->i", [A, B, v])
    tn.squeeze()
    eq, ops = tn.generate_expression()
    assert eq == "j,k,k->" and [tuple(o.shape) for o in ops] == [(1,), (9,), (9,)]
    assert "i" not in tn.dims
    assert torch.allclose(torch.einsum(eq, *ops).reshape_as(truth), truth)

    A, v, B = torch.rand(5, 9, 8), torch.rand(5), torch.rand(5, 9, 8)
    for rhs in ("ijk", "i"):
        truth = torch.einsum(f"ijk,i,ijk->{rhs}", A, v, B)
        tn = TensorNetwork(f"ijk,i,ijk->{rhs}", [A, v, B])
        tn.merge_parallel()
        eq, ops = tn.generate_expression()
        assert eq == ("ij,i,ij->ij" if rhs == "ijk" else "ij,i,ij->i")
        assert ops[0].shape == (5, 72) and ops[2].shape == (5, 72)
        assert torch.allclose(torch.einsum(eq, *ops).reshape_as(truth), truth)


def _expr(op, case):
    kw = case["kwargs"]
    if op == "forward":
        x, w, _ = C.draw_forward(case)
        return convNd_forward.einsum_expression(x, w, **kw)
    if op in ("input_vjp", "weight_vjp"):
        w, x, v = C.draw_vjp(case, forward_out_shape(case))
        if op == "input_vjp":
            return convNd_input_vjp.einsum_expression(w, v, tuple(x.shape[2:]), **kw)
        return convNd_weight_vjp.einsum_expression(x, v, tuple(w.shape[2:]), **kw)
    x = C.draw_input(case)
    mod = {"unfold": convNd_unfold, "kfc": convNd_kfc, "kfac_reduce": convNd_kfac_reduce,
           "unfold_transpose": conv_transposeNd_unfold,
           "kfac_reduce_transpose": conv_transposeNd_kfac_reduce}[op]
    return mod.einsum_expression(x, case["kernel_size"], **kw)


EQ_CASES = (
    [("forward", c) for c in C.CONV_CASES + C.SPECIAL_CASES + C.EDGE_CASES]
    + [(op, c) for op in ("input_vjp", "weight_vjp") for c in C.VJP_CASES + C.SPECIAL_CASES]
    + [("unfold", c) for c in C.UNFOLD_CASES]
    + [(op, c) for op in ("kfc", "kfac_reduce") for c in C.FACTOR_CASES]
    + [("unfold_transpose", c) for c in C.TRANSPOSE_UNFOLD_CASES]
    + [("kfac_reduce_transpose", c) for c in C.TRANSPOSE_FACTOR_CASES]
)


@pytest.mark.parametrize("op,case", EQ_CASES, ids=[f"{op}:{c['id']}" for op, c in EQ_CASES])
def test_simplified_equation_and_operand_shapes_match_reference(op, case):
    g = golden(op)
    if not g.has(case["id"]):
        pytest.skip("no golden entry (reference could not run this case)")
    ref_eq, ref_shapes = g.equation(case["id"])
    eq, ops, shape = _expr(op, case)
    assert str(eq) == ref_eq
    assert [list(o.shape) for o in ops] == ref_shapes
    assert tuple(shape) == tuple(g.data[f"{case['id']}.shape"])
    assert eq.plan.kind == op


NUM_CASES = [("forward", c) for c in C.SPECIAL_CASES] + [("unfold", C.UNFOLD_CASES[11]),
                                                         ("kfc", C.FACTOR_CASES[8]),
                                                         ("kfac_reduce", C.FACTOR_CASES[11]),
                                                         ("weight_vjp", C.VJP_CASES[9]),
                                                         ("input_vjp", C.VJP_CASES[11]),
                                                         ("unfold_transpose", C.TRANSPOSE_UNFOLD_CASES[5]),
                                                         ("kfac_reduce_transpose", C.TRANSPOSE_FACTOR_CASES[8])]


@pytest.mark.parametrize("op,case", NUM_CASES, ids=[f"{op}:{c['id']}" for op, c in NUM_CASES])
def test_simplified_expression_is_a_valid_einsum(op, case):
    """The triple stays a valid einsum: contracting the written-out operands with torch on CPU
    (explicitly, outside the product path) reproduces the reference's numbers."""
    eq, ops, shape = _expr(op, case)
    out = torch.einsum(str(eq), *dense(ops)).reshape(shape)
    if op == "forward" and case.get("bias"):
        _, _, b = C.draw_forward(case)
        out = out + b.reshape(1, -1, *([1] * (out.dim() - 2)))
    rtol = 5e-5 if "kf" in op else 2e-5
    golden(op).check(case["id"], out.numpy(), rtol, 1e-5, exact=(op in ("unfold", "unfold_transpose")))


# ---- no CPU fallback ------------------------------------------------------------------------------
def test_product_paths_refuse_cpu_tensors():
    x, w = torch.rand(1, 2, 8, 8), torch.rand(3, 2, 3, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        einconv_b200.functionals.convNd(x, w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        einconv_b200.functionals.unfoldNd(x, 3)
    eq, ops, shape = convNd_forward.einsum_expression(x, w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch.einsum(eq, *ops)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        einconv_b200.evaluator.kfc(x, 3)


def test_conv_argument_errors_match_reference():
    from einconv_b200.functionals.conv import _check_args

    x = torch.rand(1, 4, 8, 8)
    with pytest.raises(ValueError, match="kernel must be 4d"):
        _check_args(x, torch.rand(4, 4, 3))
    with pytest.raises(ValueError, match="must divide out_channels"):
        _check_args(x, torch.rand(3, 2, 3, 3), groups=2)
    with pytest.raises(ValueError, match="must equal in_channels"):
        _check_args(x, torch.rand(4, 3, 3, 3), groups=2)
    with pytest.raises(ValueError, match="Bias should have shape"):
        _check_args(x, torch.rand(4, 4, 3, 3), bias=torch.rand(5))


# ---- modules --------------------------------------------------------------------------------------
def test_convnd_module_init_matches_torch():
    torch.manual_seed(0)
    ref = nn.Conv2d(4, 6, (3, 2), stride=2, padding=1, dilation=1, groups=2)
    torch.manual_seed(0)
    mine = ConvNd(2, 4, 6, (3, 2), stride=2, padding=1, dilation=1, groups=2)
    assert torch.equal(ref.weight, mine.weight) and torch.equal(ref.bias, mine.bias)
    assert list(mine.state_dict()) == ["weight", "bias"]
    conv = ConvNd.from_nn_Conv(nn.Conv3d(2, 4, 3, padding="same", bias=False), simplify=False)
    assert conv.N == 3 and conv.bias is None and conv.padding == "same" and conv.simplify is False
    assert "simplify=False" in repr(conv) and "bias=False" in repr(conv)
    with pytest.raises(NotImplementedError):
        ConvNd(2, 4, 4, 3, padding_mode="reflect")
    with pytest.raises(ValueError):
        ConvNd(2, 4, 4, 3, groups=0)
    with pytest.raises(ValueError):
        ConvNd(2, 3, 4, 3, groups=2)
    with pytest.raises(ValueError):
        ConvNd(2, 4, 3, 3, groups=2)
    assert isinstance(UnfoldNd(3, dilation=2), nn.Module)


def test_geometry_struct_roundtrip():
    from einconv_b200 import _native as nat
    from einconv_b200.utils import resolve_geometry

    t_in, t_k, t_s, t_d, p_left, out = resolve_geometry((9, 8), 4, 1, "same", 1)
    assert p_left == (1, 1) and out == (9, 8)
    g = nat.make_geom(2, 1, 3, 4, t_in, t_k, t_s, t_d, p_left, out)
    assert g.n_dims == 2 and list(g.out)[:2] == [9, 8] and list(g.pad_left)[:2] == [1, 1]


# =============================================================================================
# call-plan cache (tensor-free planning, cached per geometry)
# =============================================================================================
def test_call_plan_cache_hits_and_keys():
    """Same (op, geometry, dtype, math, kernel family) -> the same plan object; anything else -> a new one.
    Needs the shared library (for the workspace query) but no GPU."""
    from einconv_b200 import _native as nat
    from einconv_b200 import evaluator

    evaluator.clear_plan_cache()
    args = (nat.OP_CONV_FWD, 4, 1, 8, 8, (12, 12), (3, 3), 1, 1, 1, torch.float32)
    a = evaluator.call_plan(*args)
    b = evaluator.call_plan(*args)
    assert a is b and evaluator.plan_cache_info() == {"hits": 1, "misses": 1, "size": 1}
    assert tuple(a.out) == (12, 12) and a.geom.n_dims == 2 and a.ws_bytes >= 0
    # list-valued hyper-parameters hash like tuples; 'same' padding is a different key with the same geometry
    c = evaluator.call_plan(nat.OP_CONV_FWD, 4, 1, 8, 8, (12, 12), [3, 3], 1, 1, 1, torch.float32)
    assert c is a
    d = evaluator.call_plan(nat.OP_CONV_FWD, 4, 1, 8, 8, (12, 12), (3, 3), 1, "same", 1, torch.float32)
    assert d is not a and tuple(d.out) == (12, 12)
    e = evaluator.call_plan(nat.OP_CONV_FWD, 4, 1, 8, 8, (12, 12), (3, 3), 1, 1, 1, torch.float32, False)
    assert e is not a and e.kern == nat.KERNELS_GENERIC
    # the math mode is part of the key
    evaluator.set_fp32_math("tf32")
    try:
        f = evaluator.call_plan(*args)
        assert f is not a and f.math == nat.MATH_TF32
    finally:
        evaluator.set_fp32_math("auto")
    # invalid geometries raise every time and are never cached
    size = evaluator.plan_cache_info()["size"]
    for _ in range(2):
        with pytest.raises(RuntimeError):
            evaluator.call_plan(nat.OP_CONV_FWD, 1, 1, 1, 1, (2, 2), (5, 5), 1, 0, 1, torch.float32)
    assert evaluator.plan_cache_info()["size"] == size
    assert isinstance(a.kernel, str) and a.kernel
    evaluator.clear_plan_cache()


def test_bench_profiles_exist_and_parse():
    """Every ncu summary bench.py cites for `roofline.traffic` is committed under profiles/ and yields a positive
    byte count (read + write of every kernel in the file)."""
    import importlib.util
    import pathlib

    root = pathlib.Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("bench_for_test", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    seen = 0
    for name, cls in bench.WORKLOADS.items():
        wl = cls() if name != "factors" else cls(1, 0)
        assert wl.ncu_profiles, name
        for rel in wl.ncu_profiles:
            assert (root / rel).exists(), rel
        t = bench.ncu_traffic(wl)
        assert t is not None and t["bytes"] > 0, name
        seen += 1
    assert seen == 5
