"""pytest configuration: the ``gpu`` marker and shared fixtures."""

import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Read-only view of one ``tests/golden/<op>.npz`` fixture."""

    def __init__(self, op: str):
        self.op = op
        self.data = np.load(GOLDEN / f"{op}.npz")

    def has(self, cid: str) -> bool:
        return f"{cid}.shape" in self.data

    def check(self, cid: str, result, rtol: float, atol: float, exact: bool = False):
        """Compare a full result (array-like) with the stored sample and checksums."""
        d = self.data
        res = np.asarray(result)
        assert tuple(res.shape) == tuple(d[f"{cid}.shape"]), (res.shape, d[f"{cid}.shape"])
        flat = res.astype(np.float64).reshape(-1)
        step = int(d[f"{cid}.step"])
        sample = d[f"{cid}.sample"].astype(np.float64)
        got = flat[::step]
        if exact:
            assert np.array_equal(got, sample), f"{self.op}:{cid}: sampled values differ"
            assert float(flat.sum()) == float(d[f"{cid}.sum"])
        else:
            np.testing.assert_allclose(got, sample, rtol=rtol, atol=atol, err_msg=f"{self.op}:{cid}")
            scale = max(1.0, float(np.abs(flat).sum()))
            assert abs(float(flat.sum()) - float(d[f"{cid}.sum"])) <= 10 * rtol * scale + atol
            ssq = float(d[f"{cid}.sumsq"])
            assert abs(float((flat * flat).sum()) - ssq) <= 20 * rtol * max(1.0, ssq) + atol

    def equation(self, cid: str):
        import json

        return str(self.data[f"{cid}.equation"]), json.loads(str(self.data[f"{cid}.opshapes"]))

    def insum(self, cid: str) -> float:
        return float(self.data[f"{cid}.insum"])


_cache = {}


def golden(op: str) -> Golden:
    if op not in _cache:
        _cache[op] = Golden(op)
    return _cache[op]


def forward_out_shape(case):
    """Output shape of a forward case from the oracle's geometry (independent of the product)."""
    from oracle import direct

    x, w, kw = case["x"], case["w"], case["kwargs"]
    g = direct.geometry(x[0], kw.get("groups", 1), w[1], w[0] // kw.get("groups", 1), x[2:], w[2:],
                        kw.get("stride", 1), kw.get("padding", 0), kw.get("dilation", 1))
    return (x[0], w[0], *[g.out[d] for d in range(len(x) - 2)])


@pytest.fixture(autouse=True)
def _fresh_plan_cache():
    """Call plans cache the workspace size of the route chosen when they were made; tests that flip the
    library's diagnostic EINCONV_* switches (monkeypatch.setenv) must not see a plan made under another
    setting."""
    from einconv_b200 import evaluator

    evaluator.clear_plan_cache()
    yield
