"""Golden vectors for the transpose-convolution networks, produced by running
the REFERENCE itself on CPU — same conventions and record format as ``make_golden.py``:

    PYTHONPATH=/root/reference:/root/repo python tests/golden/make_golden_transpose.py

Writes ``unfold_transpose.npz`` and ``kfac_reduce_transpose.npz`` next to this file.
"""

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

import einconv  # noqa: E402  (the reference, from PYTHONPATH=/root/reference)
from einconv import functionals  # noqa: E402
from einconv.expressions import conv_transposeNd_kfac_reduce, conv_transposeNd_unfold  # noqa: E402

from tests import cases as C  # noqa: E402
from tests.golden.make_golden import record  # noqa: E402

assert "/root/reference" in einconv.__file__, einconv.__file__


def main():
    torch.set_num_threads(8)
    store = {}
    for case in C.TRANSPOSE_UNFOLD_CASES:
        x = C.draw_input(case)
        u = functionals.unfoldNd_transpose(x, case["kernel_size"], **case["kwargs"], simplify=True)
        eq, ops, _ = conv_transposeNd_unfold.einsum_expression(x, case["kernel_size"], **case["kwargs"], simplify=True)
        record(store, case["id"], u, [x], eq, ops)
        print("unfold_transpose", case["id"], tuple(u.shape), flush=True)
    np.savez_compressed(HERE / "unfold_transpose.npz", **store)

    store = {}
    for case in C.TRANSPOSE_FACTOR_CASES:
        x = C.draw_input(case)
        eq, ops, shape = conv_transposeNd_kfac_reduce.einsum_expression(
            x, case["kernel_size"], **case["kwargs"], simplify=True
        )
        res = torch.einsum(eq, *ops).reshape(shape)
        record(store, case["id"], res, [x], eq, ops)
        print("kfac_reduce_transpose", case["id"], tuple(res.shape), flush=True)
    np.savez_compressed(HERE / "kfac_reduce_transpose.npz", **store)


if __name__ == "__main__":
    main()
