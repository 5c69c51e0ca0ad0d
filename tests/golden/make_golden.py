"""Generate golden vectors by running the REFERENCE itself (f-dangel/einconv) on CPU.

Run in the build container, where the reference is mounted read-only:

    PYTHONPATH=/root/reference:/root/repo python tests/golden/make_golden.py

The reference is pure Python (torch + einops), so it is imported directly; its sources are never
copied.  For every case of ``tests/cases.py`` the reference's public API is called exactly as its
own tests call it (``functionals.convNd`` / ``unfoldNd``, ``expressions.*.einsum_expression`` +
``torch.einsum``), and the result is stored as a small fixture:

* ``<id>.shape``   result shape
* ``<id>.sample``  ``result.flatten()[::step]`` with ``step`` chosen so that <= 2048 values are kept
* ``<id>.step``    that step
* ``<id>.sum`` / ``<id>.sumsq``  float64 checksums over the FULL result
* ``<id>.insum``   float64 sum of all drawn inputs (guards against RNG drift on another machine)
* ``<id>.equation`` / ``<id>.opshapes``  the reference's simplified equation and operand shapes

Cases the reference cannot finish within ``TIME_LIMIT`` seconds or ``MEM_LIMIT`` bytes of
intermediates on this box are recorded in ``skipped.json`` with the reason.
"""

import json
import signal
import sys
import time
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

import einconv  # noqa: E402  (the reference, from PYTHONPATH=/root/reference)
from einconv import functionals  # noqa: E402
from einconv.expressions import (  # noqa: E402
    convNd_forward,
    convNd_input_vjp,
    convNd_kfac_reduce,
    convNd_kfc,
    convNd_unfold,
    convNd_weight_vjp,
)

from tests import cases as C  # noqa: E402

assert "/root/reference" in einconv.__file__, einconv.__file__

TIME_LIMIT = 120
MAX_SAMPLE = 2048


class Timeout(Exception):
    pass


def _alarm(signum, frame):
    raise Timeout()


signal.signal(signal.SIGALRM, _alarm)


def record(store, cid, result, inputs, equation=None, operands=None):
    flat = result.detach().to(torch.float64).flatten().numpy()
    step = max(1, -(-flat.size // MAX_SAMPLE))
    store[f"{cid}.shape"] = np.asarray(result.shape, dtype=np.int64)
    store[f"{cid}.sample"] = flat[::step].astype(np.float64 if result.dtype == torch.float64 else np.float32)
    store[f"{cid}.step"] = np.asarray(step)
    store[f"{cid}.sum"] = np.asarray(flat.sum())
    store[f"{cid}.sumsq"] = np.asarray((flat * flat).sum())
    store[f"{cid}.insum"] = np.asarray(sum(float(t.double().sum()) for t in inputs if t is not None))
    if equation is not None:
        store[f"{cid}.equation"] = np.asarray(str(equation))
        store[f"{cid}.opshapes"] = np.asarray(json.dumps([list(op.shape) for op in operands]))


def guarded(skipped, op, cid, fn):
    t0 = time.time()
    signal.alarm(TIME_LIMIT)
    try:
        fn()
        print(f"  {op:12s} {cid:70s} {time.time() - t0:6.1f}s", flush=True)
    except Timeout:
        skipped.append(dict(op=op, id=cid, reason=f"reference exceeded {TIME_LIMIT}s on CPU"))
        print(f"  {op:12s} {cid:70s} TIMEOUT", flush=True)
    except (RuntimeError, MemoryError, ValueError) as e:
        skipped.append(dict(op=op, id=cid, reason=f"{type(e).__name__}: {str(e)[:160]}"))
        print(f"  {op:12s} {cid:70s} {type(e).__name__}: {str(e)[:80]}", flush=True)
    finally:
        signal.alarm(0)


def main():
    torch.set_num_threads(8)
    skipped = []

    # ---- forward ------------------------------------------------------------------------
    store = {}
    for case in C.CONV_CASES + C.SPECIAL_CASES + C.EDGE_CASES:
        def run(case=case):
            x, w, b = C.draw_forward(case)
            y = functionals.convNd(x, w, bias=b, **case["kwargs"], simplify=True)
            eq, ops, _ = convNd_forward.einsum_expression(x, w, **case["kwargs"], simplify=True)
            record(store, case["id"], y, [x, w, b], eq, ops)
        guarded(skipped, "forward", case["id"], run)
    np.savez_compressed(HERE / "forward.npz", **store)

    # ---- VJPs ---------------------------------------------------------------------------
    store_i, store_w = {}, {}
    for case in C.VJP_CASES + C.SPECIAL_CASES:
        def shapes(case=case):
            x0 = torch.zeros(*case["x"])
            w0 = torch.zeros(*case["w"])
            eq, ops, shape = convNd_forward.einsum_expression(x0, w0, **case["kwargs"], simplify=False)
            return shape

        def run_i(case=case):
            w, x, v = C.draw_vjp(case, shapes())
            eq, ops, shape = convNd_input_vjp.einsum_expression(
                w, v, tuple(x.shape[2:]), **case["kwargs"], simplify=True
            )
            record(store_i, case["id"], torch.einsum(eq, *ops).reshape(shape), [w, x, v], eq, ops)

        def run_w(case=case):
            w, x, v = C.draw_vjp(case, shapes())
            eq, ops, shape = convNd_weight_vjp.einsum_expression(
                x, v, tuple(w.shape[2:]), **case["kwargs"], simplify=True
            )
            record(store_w, case["id"], torch.einsum(eq, *ops).reshape(shape), [w, x, v], eq, ops)

        guarded(skipped, "input_vjp", case["id"], run_i)
        guarded(skipped, "weight_vjp", case["id"], run_w)
    np.savez_compressed(HERE / "input_vjp.npz", **store_i)
    np.savez_compressed(HERE / "weight_vjp.npz", **store_w)

    # ---- unfold -------------------------------------------------------------------------
    store = {}
    for case in C.UNFOLD_CASES:
        def run(case=case):
            x = C.draw_input(case)
            u = functionals.unfoldNd(x, case["kernel_size"], **case["kwargs"], simplify=True)
            eq, ops, _ = convNd_unfold.einsum_expression(x, case["kernel_size"], **case["kwargs"], simplify=True)
            record(store, case["id"], u, [x], eq, ops)
        guarded(skipped, "unfold", case["id"], run)
    np.savez_compressed(HERE / "unfold.npz", **store)

    # ---- Kronecker factors ----------------------------------------------------------------
    store_k, store_r = {}, {}
    for case in C.FACTOR_CASES:
        ks = case["kernel_size"]
        k_vol = ks ** (len(case["x"]) - 2) if isinstance(ks, int) else int(np.prod(ks))
        if len(case["x"]) > 4 and k_vol > 30:
            # 3-d factors with 5x5x5-sized kernels: the reference's left-to-right einsum asks for
            # more than 40 GB of intermediates here (measured: DefaultCPUAllocator failure under
            # `ulimit -v 40000000`).  These cases are checked against the C oracle.
            skipped.append(dict(op="kfc/kfac_reduce", id=case["id"],
                                reason="reference einsum needs > 40 GB of intermediates on CPU"))
            continue

        def run_k(case=case):
            x = C.draw_input(case)
            eq, ops, shape = convNd_kfc.einsum_expression(x, case["kernel_size"], **case["kwargs"], simplify=True)
            record(store_k, case["id"], torch.einsum(eq, *ops).reshape(shape), [x], eq, ops)

        def run_r(case=case):
            x = C.draw_input(case)
            eq, ops, shape = convNd_kfac_reduce.einsum_expression(
                x, case["kernel_size"], **case["kwargs"], simplify=True
            )
            record(store_r, case["id"], torch.einsum(eq, *ops).reshape(shape), [x], eq, ops)

        guarded(skipped, "kfc", case["id"], run_k)
        guarded(skipped, "kfac_reduce", case["id"], run_r)
    np.savez_compressed(HERE / "kfc.npz", **store_k)
    np.savez_compressed(HERE / "kfac_reduce.npz", **store_r)

    # ---- index patterns (dense, exact) ------------------------------------------------------
    store = {}
    for n, kw in enumerate(C.INDEX_PATTERN_CASES):
        pat = einconv.index_pattern(**kw)
        store[f"{n}.pattern"] = np.packbits(pat.numpy())
        store[f"{n}.shape"] = np.asarray(pat.shape)
    np.savez_compressed(HERE / "index_pattern.npz", **store)

    (HERE / "skipped.json").write_text(json.dumps(skipped, indent=1))
    meta = dict(torch=torch.__version__, reference=str(Path(einconv.__file__).parent),
                threads=torch.get_num_threads(), generated=time.strftime("%Y-%m-%d"))
    (HERE / "meta.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
