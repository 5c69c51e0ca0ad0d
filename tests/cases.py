"""Parity cases: the reference's own test-case data, restated as plain tuples.

Shapes, seeds and hyper-parameters come from the reference's case lists (they are data):
``test/functionals/conv_cases.py``, ``test/modules/conv_cases.py`` (N = 4, 5, 6),
``test/functionals/unfold_cases.py``, ``test/expressions/convNd_{input_vjp,weight_vjp,kfc,
kfac_reduce}_cases.py`` and ``test/conv_index_pattern_cases.py``.  Inputs are ``torch.rand`` after
``torch.manual_seed(seed)``, drawn in the order the reference's tests draw them (x, weight, bias
for forward; weight, x, v for the VJPs).

Each case is a dict: ``id``, ``seed``, ``x`` (shape), and op-specific keys.  ``BASELINE`` holds the
BASELINE.json configurations at reduced batch (for oracle-sized parity) and ``EDGE`` adds the edge
cases of SURVEY.md App. D (empty batch, padding wider than the kernel, asymmetric 'same', ...).
"""

from typing import Any, Dict, List

import torch


def _id(prefix: str, case: Dict[str, Any]) -> str:
    parts = [prefix, "x".join(map(str, case["x"]))]
    for key in ("w", "kernel_size"):
        if key in case:
            v = case[key]
            parts.append(f"{key[0]}" + ("x".join(map(str, v)) if isinstance(v, tuple) else str(v)))
    for k, v in case.get("kwargs", {}).items():
        if isinstance(v, tuple):
            v = "x".join(map(str, v))
        parts.append(f"{k[:3]}{v}")
    if case.get("bias"):
        parts.append("bias")
    if case.get("dtype", torch.float32) != torch.float32:
        parts.append(str(case["dtype"]).replace("torch.", ""))
    return "-".join(parts)


def _mk(prefix: str, cases: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    for c in cases:
        c.setdefault("seed", 0)
        c.setdefault("kwargs", {})
        c.setdefault("id", _id(prefix, c))
    return cases


# -------------------------------------------------------------------------------------------
# forward (functional) — test/functionals/conv_cases.py:7-234 and modules N = 4..6
# -------------------------------------------------------------------------------------------
CONV_CASES = _mk(
    "fwd",
    [
        # 1d
        dict(x=(2, 3, 50), w=(4, 3, 5)),
        dict(x=(2, 3, 50), w=(4, 3, 5), bias=True),
        dict(x=(2, 4, 50), w=(6, 2, 5), bias=True, kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 50), w=(4, 3, 5), bias=True, kwargs=dict(padding=2)),
        dict(x=(2, 3, 50), w=(4, 3, 5), bias=True, kwargs=dict(padding=(2,), stride=(3,), dilation=(1,))),
        dict(x=(2, 3, 50), w=(4, 3, 5), kwargs=dict(dilation=2)),
        dict(x=(2, 3, 20), w=(4, 3, 5), bias=True, kwargs=dict(padding="same")),
        # 2d
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 5)),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 5), bias=True),
        dict(x=(2, 4, 50, 40), w=(6, 2, 5, 5), bias=True, kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 6), bias=True, kwargs=dict(padding=2)),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 3), bias=True,
             kwargs=dict(padding=(2, 1), stride=(3, 2), dilation=(1, 2))),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 6), kwargs=dict(dilation=2)),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 3), bias=True, kwargs=dict(padding="same", dilation=(1, 2))),
        # 3d
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 5, 5)),
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 5, 5), bias=True),
        dict(x=(2, 4, 50, 40, 30), w=(6, 2, 5, 5, 5), bias=True, kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 6, 7), bias=True, kwargs=dict(padding=2)),
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 3, 2), bias=True,
             kwargs=dict(padding=(2, 1, 0), stride=(3, 2, 1), dilation=(1, 2, 3))),
        dict(x=(2, 3, 50, 40, 40), w=(4, 3, 5, 6, 7), kwargs=dict(dilation=2)),
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 3, 2), bias=True,
             kwargs=dict(padding="same", dilation=(1, 2, 2))),
        # module cases with groups / tuples (test/modules/conv_cases.py), 1d-3d
        dict(x=(2, 3, 50), w=(9, 1, 5), bias=True, kwargs=dict(stride=2, padding=1, dilation=2, groups=3)),
        dict(x=(2, 6, 71), w=(9, 2, 5), bias=True, kwargs=dict(padding="same", dilation=(2,), groups=3)),
        dict(x=(2, 6, 99, 88), w=(9, 2, 5, 4), bias=True,
             kwargs=dict(stride=(2, 3), padding=(1, 0), dilation=(2, 1), groups=3)),
        dict(x=(2, 6, 63, 52), w=(9, 2, 5, 4), bias=True,
             kwargs=dict(padding="same", dilation=(1, 2), groups=3)),
        dict(x=(2, 6, 37, 41, 28), w=(9, 2, 5, 4, 3), bias=True,
             kwargs=dict(padding="same", dilation=(2, 1, 3), groups=3)),
        # N = 4, 5, 6 (the reference checks these against JAX)
        dict(x=(2, 3, 25, 20, 15, 10), w=(4, 3, 5, 5, 5, 5), bias=True),
        dict(x=(2, 2, 40, 35, 30, 25), w=(4, 1, 5, 5, 4, 3), bias=True,
             kwargs=dict(stride=(2, 3, 2, 1), padding=(2, 0, 2, 1), dilation=(2, 1, 3, 2), groups=2)),
        dict(x=(2, 2, 21, 26, 33, 17), w=(4, 1, 5, 5, 4, 3), bias=True,
             kwargs=dict(padding="same", dilation=(2, 1, 3, 2), groups=2)),
        dict(x=(2, 3, 15, 10, 8, 6, 5), w=(4, 3, 3, 3, 3, 3, 3)),
        dict(x=(2, 2, 20, 15, 10, 15, 20), w=(4, 1, 5, 4, 2, 3, 4),
             kwargs=dict(stride=(4, 3, 1, 2, 4), padding=(2, 0, 1, 1, 3), groups=2)),
        dict(x=(2, 2, 10, 10, 10, 10, 5, 5), w=(4, 1, 3, 2, 3, 4, 2, 2), bias=True,
             kwargs=dict(stride=(3, 2, 2, 3, 3, 2), padding=(2, 0, 0, 2, 1, 1), groups=2)),
    ],
)

# -------------------------------------------------------------------------------------------
# VJPs — test/expressions/convNd_{input,weight}_vjp_cases.py (identical lists)
# -------------------------------------------------------------------------------------------
VJP_CASES = _mk(
    "vjp",
    [
        dict(x=(2, 3, 50), w=(4, 3, 5)),
        dict(x=(2, 3, 50), w=(4, 3, 5), kwargs=dict(stride=2)),
        dict(x=(2, 4, 50), w=(6, 2, 5), kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 50), w=(4, 3, 5), kwargs=dict(padding=2)),
        dict(x=(2, 3, 50), w=(4, 3, 5), kwargs=dict(padding=(2,), stride=(3,), dilation=(1,))),
        dict(x=(2, 3, 50), w=(4, 3, 5), kwargs=dict(dilation=2)),
        dict(x=(2, 3, 20), w=(4, 3, 5), kwargs=dict(padding="same")),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 5)),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 5), kwargs=dict(stride=2)),
        dict(x=(2, 4, 50, 40), w=(6, 2, 5, 5), kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 6), kwargs=dict(padding=2)),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 3), kwargs=dict(padding=(2, 1), stride=(3, 2), dilation=(1, 2))),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 6), kwargs=dict(dilation=2)),
        dict(x=(2, 3, 50, 40), w=(4, 3, 5, 3), kwargs=dict(padding="same", dilation=(1, 2))),
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 5, 5)),
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 5, 5), kwargs=dict(stride=2)),
        dict(x=(2, 4, 50, 40, 30), w=(6, 2, 5, 5, 5), kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 6, 7), kwargs=dict(padding=2)),
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 3, 2),
             kwargs=dict(padding=(2, 1, 0), stride=(3, 2, 1), dilation=(1, 2, 3))),
        dict(x=(2, 3, 50, 40, 40), w=(4, 3, 5, 6, 7), kwargs=dict(dilation=2)),
        dict(x=(2, 3, 50, 40, 30), w=(4, 3, 5, 3, 2), kwargs=dict(padding="same", dilation=(1, 2, 2))),
    ],
)

# -------------------------------------------------------------------------------------------
# unfold — test/functionals/unfold_cases.py:11-138 (incl. the float64 regression :89-95)
# -------------------------------------------------------------------------------------------
UNFOLD_CASES = _mk(
    "unf",
    [
        dict(seed=0, x=(2, 3, 50), kernel_size=1),
        dict(seed=1, x=(2, 3, 50), kernel_size=2),
        dict(seed=2, x=(2, 3, 50), kernel_size=3),
        dict(seed=3, x=(2, 3, 50), kernel_size=3, kwargs=dict(dilation=2)),
        dict(seed=4, x=(2, 3, 50), kernel_size=3, kwargs=dict(dilation=2, padding=1)),
        dict(seed=5, x=(2, 3, 50), kernel_size=3, kwargs=dict(dilation=2, padding=1, stride=2)),
        dict(seed=0, x=(2, 3, 50, 40), kernel_size=1),
        dict(seed=1, x=(2, 3, 50, 40), kernel_size=2),
        dict(seed=2, x=(2, 3, 50, 40), kernel_size=(3, 2)),
        dict(seed=3, x=(2, 3, 50, 40), kernel_size=(3, 2), kwargs=dict(dilation=2)),
        dict(seed=4, x=(2, 3, 50, 40), kernel_size=(3, 2), kwargs=dict(dilation=2, padding=1)),
        dict(seed=5, x=(2, 3, 50, 40), kernel_size=(3, 2), kwargs=dict(dilation=2, padding=1, stride=2)),
        dict(seed=0, x=(2, 3, 50, 40), kernel_size=1, dtype=torch.float64),
        dict(seed=0, x=(2, 3, 50, 40, 30), kernel_size=1),
        dict(seed=1, x=(2, 3, 50, 40, 30), kernel_size=2),
        dict(seed=2, x=(2, 3, 50, 40, 30), kernel_size=(4, 3, 2)),
        dict(seed=3, x=(2, 3, 50, 40, 30), kernel_size=(4, 3, 2), kwargs=dict(dilation=2)),
        dict(seed=4, x=(2, 3, 50, 40, 30), kernel_size=(4, 3, 2), kwargs=dict(dilation=2, padding=1)),
        dict(seed=5, x=(2, 3, 50, 40, 30), kernel_size=(4, 3, 2),
             kwargs=dict(dilation=2, padding=1, stride=2)),
    ],
)

# -------------------------------------------------------------------------------------------
# Kronecker factors — test/expressions/convNd_{kfc,kfac_reduce}_cases.py (identical lists)
# -------------------------------------------------------------------------------------------
FACTOR_CASES = _mk(
    "fac",
    [
        dict(x=(2, 3, 50), kernel_size=5),
        dict(x=(2, 3, 50), kernel_size=5, kwargs=dict(stride=2)),
        dict(x=(2, 4, 50), kernel_size=5, kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 50), kernel_size=5, kwargs=dict(padding=2)),
        dict(x=(2, 3, 50), kernel_size=(5,), kwargs=dict(padding=(2,), stride=(3,), dilation=(1,))),
        dict(x=(2, 3, 50), kernel_size=5, kwargs=dict(dilation=2)),
        dict(x=(2, 3, 20), kernel_size=5, kwargs=dict(padding="same")),
        dict(x=(2, 3, 50, 40), kernel_size=5),
        dict(x=(2, 3, 50, 40), kernel_size=5, kwargs=dict(stride=2)),
        dict(x=(2, 4, 50, 40), kernel_size=5, kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 50, 40), kernel_size=6, kwargs=dict(padding=2)),
        dict(x=(2, 3, 50, 40), kernel_size=(5, 3), kwargs=dict(padding=(2, 1), stride=(3, 2), dilation=(1, 2))),
        dict(x=(2, 3, 50, 40), kernel_size=(5, 6), kwargs=dict(dilation=2)),
        dict(x=(2, 3, 50, 40), kernel_size=(5, 3), kwargs=dict(padding="same", dilation=(1, 2))),
        dict(x=(2, 3, 50, 40, 30), kernel_size=5),
        dict(x=(2, 3, 50, 40, 30), kernel_size=5, kwargs=dict(stride=2)),
        dict(x=(2, 4, 50, 40, 30), kernel_size=5, kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 50, 40, 30), kernel_size=(5, 6, 7), kwargs=dict(padding=2)),
        dict(x=(2, 3, 50, 40, 30), kernel_size=(5, 3, 2),
             kwargs=dict(padding=(2, 1, 0), stride=(3, 2, 1), dilation=(1, 2, 3))),
        dict(x=(2, 3, 50, 40, 40), kernel_size=(5, 6, 7), kwargs=dict(dilation=2)),
        dict(x=(2, 3, 50, 40, 30), kernel_size=(5, 3, 2), kwargs=dict(padding="same", dilation=(1, 2, 2))),
    ],
)

# -------------------------------------------------------------------------------------------
# transpose convolutions.  x has the associated convolution's OUTPUT sizes.
# test/functionals/transpose_unfold_cases.py:7-72 and
# test/expressions/conv_transposeNd_kfac_reduce_cases.py:7-125, plus a few larger / N = 4 cases
# -------------------------------------------------------------------------------------------
TRANSPOSE_UNFOLD_CASES = _mk(
    "tunf",
    [
        dict(x=(2, 3, 50), kernel_size=1),
        dict(x=(5, 2, 9), kernel_size=3, kwargs=dict(stride=2, padding=1, output_padding=1, dilation=2)),
        dict(x=(2, 3, 10, 10), kernel_size=1),
        dict(x=(5, 2, 9, 7), kernel_size=(3, 2),
             kwargs=dict(stride=(2, 1), padding=(1, 0), output_padding=(1, 0), dilation=(2, 2))),
        dict(x=(2, 3, 10, 10, 10), kernel_size=1),
        dict(x=(5, 2, 9, 7, 6), kernel_size=(3, 2, 4),
             kwargs=dict(stride=(2, 1, 3), padding=(1, 0, 2), output_padding=(1, 0, 2), dilation=(2, 2, 1))),
        # additions
        dict(x=(3, 4, 17, 13), kernel_size=3, kwargs=dict(padding=1)),
        dict(x=(2, 3, 12, 11), kernel_size=(4, 3), kwargs=dict(stride=2, padding=1)),
        dict(x=(1, 2, 4, 5, 3, 4), kernel_size=(2, 3, 2, 2), kwargs=dict(stride=(1, 2, 1, 2), padding=(0, 1, 0, 0))),
    ],
)

TRANSPOSE_FACTOR_CASES = _mk(
    "tfac",
    [
        dict(x=(2, 3, 8), kernel_size=5),
        dict(x=(2, 3, 8), kernel_size=5, kwargs=dict(stride=2)),
        dict(x=(2, 4, 8), kernel_size=5, kwargs=dict(stride=3, groups=2)),
        dict(x=(2, 3, 8), kernel_size=5, kwargs=dict(padding=2)),
        dict(x=(2, 3, 8), kernel_size=5, kwargs=dict(output_padding=1, stride=2)),
        dict(x=(2, 3, 8), kernel_size=5, kwargs=dict(dilation=2)),
        dict(x=(2, 3, 8), kernel_size=(3,), kwargs=dict(padding=(1,), stride=(2,), dilation=(1,), output_padding=(1,))),
        dict(x=(2, 3, 8, 7), kernel_size=5),
        dict(x=(2, 3, 8, 7), kernel_size=(5, 3),
             kwargs=dict(padding=(1, 2), stride=(2, 3), dilation=(2, 1), output_padding=(1, 2))),
        dict(x=(2, 3, 8, 7, 6), kernel_size=5),
        dict(x=(2, 3, 8, 7, 6), kernel_size=(5, 3, 2),
             kwargs=dict(padding=(0, 1, 2), stride=(3, 2, 1), dilation=(1, 2, 3), output_padding=(2, 0, 0))),
        # additions
        dict(x=(4, 12, 14, 14), kernel_size=3, kwargs=dict(padding=1, groups=3)),
        dict(x=(3, 40, 9, 9), kernel_size=3, kwargs=dict(stride=2, padding=1, output_padding=1)),
    ],
)

# -------------------------------------------------------------------------------------------
# index pattern — test/conv_index_pattern_cases.py:5-24
# -------------------------------------------------------------------------------------------
INDEX_PATTERN_CASES = [
    dict(input_size=13, kernel_size=2),
    dict(input_size=15, kernel_size=3, stride=2),
    dict(input_size=11, kernel_size=4, padding=2),
    dict(input_size=11, kernel_size=4, padding=6),
    dict(input_size=11, kernel_size=4, padding="valid"),
    dict(input_size=39, kernel_size=5, padding="same", dilation=2),
    dict(input_size=20, kernel_size=3, stride=2),
    dict(input_size=91, kernel_size=4, stride=2, padding=2, dilation=3),
    dict(input_size=91, kernel_size=4, stride=2, padding=20, dilation=3),
]

# -------------------------------------------------------------------------------------------
# simplifier specialisations: dense / 1x1 / down-sampling / depthwise
# -------------------------------------------------------------------------------------------
SPECIAL_CASES = _mk(
    "spc",
    [
        dict(x=(2, 8, 12, 12), w=(6, 8, 1, 1)),                                  # 1x1 pointwise
        dict(x=(2, 8, 12, 12), w=(6, 8, 1, 1), kwargs=dict(stride=2)),           # 1x1 stride 2 (down-sampling)
        dict(x=(2, 4, 12, 12), w=(6, 4, 3, 3), kwargs=dict(stride=3)),           # patchify K == S
        dict(x=(2, 4, 12, 16), w=(6, 4, 2, 4), kwargs=dict(stride=(2, 4))),      # patchify, mixed
        dict(x=(2, 4, 12, 12), w=(6, 4, 2, 2), kwargs=dict(stride=3)),           # S > K
        dict(x=(2, 16, 14, 14), w=(16, 1, 3, 3), bias=True, kwargs=dict(padding=1, groups=16)),   # depthwise
        dict(x=(2, 16, 15, 13), w=(32, 1, 3, 3), kwargs=dict(padding=1, stride=2, groups=16)),    # depthwise x2
        dict(x=(2, 6, 17, 11), w=(6, 1, 5, 5), kwargs=dict(padding=2, dilation=2, groups=6)),     # depthwise 5x5 dil
        dict(x=(3, 5, 33), w=(5, 1, 4), bias=True, kwargs=dict(padding=3, stride=2, groups=5)),   # depthwise 1d
        dict(x=(1, 4, 9, 9), w=(4, 1, 2, 3), kwargs=dict(groups=4)),             # depthwise generic taps
    ],
)

# -------------------------------------------------------------------------------------------
# edge cases (SURVEY.md App. D and the reference's own odd ones)
# -------------------------------------------------------------------------------------------
EDGE_CASES = _mk(
    "edge",
    [
        dict(x=(1, 1, 11), w=(1, 1, 4), kwargs=dict(padding=6)),                 # padding wider than kernel
        dict(x=(1, 2, 91), w=(3, 2, 4), kwargs=dict(stride=2, padding=20, dilation=3)),
        dict(x=(2, 3, 9, 8), w=(4, 3, 4, 4), kwargs=dict(padding="same")),       # asymmetric 'same' (1, 2)
        dict(x=(1, 1, 5, 5), w=(1, 1, 5, 5)),                                    # single output pixel
        dict(x=(3, 2, 7, 7), w=(2, 2, 7, 7), kwargs=dict(padding=3)),            # kernel == input, padded
        dict(x=(2, 3, 1, 1), w=(5, 3, 1, 1), bias=True),                         # 1x1 image
        dict(x=(70, 3, 5, 5), w=(130, 3, 3, 3), bias=True),                      # ragged GEMM tiles
    ],
)

# BASELINE.json configurations, reduced so that the CPU oracle finishes in seconds
BASELINE_SMALL = dict(
    c1=dict(x=(2, 64, 56, 56), w=(64, 64, 3, 3), kwargs=dict(padding=1)),
    c2=dict(x=(1, 4, 16, 56, 56), kernel_size=3, kwargs=dict(stride=2, dilation=2, padding=0)),
    c3=dict(x=(2, 256, 28, 28), w=(256, 1, 3, 3), kwargs=dict(padding=1, groups=256)),
    c4=dict(x=(1, 8, 6, 28, 28), w=(8, 8, 3, 3, 3), kwargs=dict(padding=1)),
    c5=dict(x=(4, 16, 14, 14), kernel_size=3, kwargs=dict(padding=1)),
)


def draw_forward(case, device="cpu"):
    """x, weight, bias in the reference's draw order (test/functionals/test_conv.py:55-64)."""
    torch.manual_seed(case["seed"])
    dtype = case.get("dtype", torch.float32)
    x = torch.rand(*case["x"]).to(dtype)
    w = torch.rand(*case["w"]).to(dtype)
    b = torch.rand(case["w"][0]).to(dtype) if case.get("bias") else None
    return x.to(device), w.to(device), None if b is None else b.to(device)


def draw_vjp(case, out_shape, device="cpu"):
    """weight, x, v in the reference's draw order (test_convNd_weight_vjp.py:36-46)."""
    torch.manual_seed(case["seed"])
    w = torch.rand(*case["w"])
    x = torch.rand(*case["x"])
    v = torch.rand(*out_shape)
    return w.to(device), x.to(device), v.to(device)


def draw_input(case, device="cpu"):
    torch.manual_seed(case["seed"])
    return torch.rand(*case["x"]).to(case.get("dtype", torch.float32)).to(device)
