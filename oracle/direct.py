"""ORACLE (test infrastructure): ctypes/numpy front-end of ``einconv_oracle.c``.

Geometry follows ``einconv/utils.py:97-160`` (restated in ``geometry`` below, independent of the
product package on purpose).
"""

import ctypes
import subprocess
from math import floor
from pathlib import Path
from typing import Sequence, Tuple, Union

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libeinconv_oracle.so"
MAXD = 8


class Geom(ctypes.Structure):
    _fields_ = [
        ("n_dims", ctypes.c_int32),
        ("groups", ctypes.c_int32),
        ("batch", ctypes.c_int64),
        ("c_in_g", ctypes.c_int64),
        ("c_out_g", ctypes.c_int64),
        ("in_", ctypes.c_int32 * MAXD),
        ("k", ctypes.c_int32 * MAXD),
        ("stride", ctypes.c_int32 * MAXD),
        ("dilation", ctypes.c_int32 * MAXD),
        ("pad_left", ctypes.c_int32 * MAXD),
        ("out", ctypes.c_int32 * MAXD),
    ]


def build() -> Path:
    """Compile the C oracle if needed (gcc, seconds)."""
    src = HERE / "einconv_oracle.c"
    if not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
    return _lib


def _tup(h, n):
    return (h,) * n if isinstance(h, int) else tuple(h)


def paddings(k: int, s: int, p: Union[int, str], d: int) -> Tuple[int, int]:
    """einconv/utils.py:97-132."""
    if isinstance(p, str):
        if p == "valid":
            return 0, 0
        if p == "same":
            if s != 1:
                raise ValueError("padding='same' is not supported for strided convolutions.")
            total = d * (k - 1)
            return total // 2, total - total // 2
        raise ValueError(f"Unknown string-value for padding: {p!r}.")
    return p, p


def output_size(i: int, k: int, s: int, p: Union[int, str], d: int) -> int:
    """einconv/utils.py:135-160."""
    pl, pr = paddings(k, s, p, d)
    span = k + (k - 1) * (d - 1)
    return 1 + floor((i + pl + pr - span) / s)


def geometry(batch, groups, c_in_g, c_out_g, input_size: Sequence[int], kernel_size, stride=1,
             padding=0, dilation=1) -> Geom:
    n = len(input_size)
    ks, ss, ds = _tup(kernel_size, n), _tup(stride, n), _tup(dilation, n)
    ps = (padding,) * n if isinstance(padding, (int, str)) else tuple(padding)
    g = Geom()
    g.n_dims, g.groups, g.batch, g.c_in_g, g.c_out_g = n, groups, batch, c_in_g, c_out_g
    for d in range(n):
        g.in_[d], g.k[d], g.stride[d], g.dilation[d] = input_size[d], ks[d], ss[d], ds[d]
        g.pad_left[d] = paddings(ks[d], ss[d], ps[d], ds[d])[0]
        g.out[d] = output_size(input_size[d], ks[d], ss[d], ps[d], ds[d])
    return g


def _out_sizes(g: Geom):
    return [g.out[d] for d in range(g.n_dims)]


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def unfold(x: np.ndarray, kernel_size, dilation=1, padding=0, stride=1) -> np.ndarray:
    """Bit-exact im2col of ``x`` [B, C, *in] (any 2/4/8-byte dtype) -> [B, C*Kt, Ot]."""
    x = np.ascontiguousarray(x)
    B, C = x.shape[:2]
    g = geometry(B, 1, C, 0, x.shape[2:], kernel_size, stride, padding, dilation)
    kt = int(np.prod([g.k[d] for d in range(g.n_dims)]))
    ot = int(np.prod(_out_sizes(g)))
    out = np.empty((B, C * kt, ot), dtype=x.dtype)
    lib().oracle_unfold(_p(x), _p(out), ctypes.c_int(x.dtype.itemsize), ctypes.byref(g))
    return out


def fold(gu, channels, input_size, kernel_size, dilation=1, padding=0, stride=1) -> np.ndarray:
    gu = _f64(gu)
    B = gu.shape[0]
    g = geometry(B, 1, channels, 0, input_size, kernel_size, stride, padding, dilation)
    gx = np.empty((B, channels, *input_size), dtype=np.float64)
    lib().oracle_fold(_p(gu), _p(gx), ctypes.byref(g))
    return gx


def conv_forward(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1) -> np.ndarray:
    x, w = _f64(x), _f64(w)
    B = x.shape[0]
    g = geometry(B, groups, w.shape[1], w.shape[0] // groups, x.shape[2:], w.shape[2:], stride,
                 padding, dilation)
    y = np.empty((B, w.shape[0], *_out_sizes(g)), dtype=np.float64)
    b = None if bias is None else _f64(bias)
    lib().oracle_conv_fwd(_p(x), _p(w), None if b is None else _p(b), _p(y), ctypes.byref(g))
    return y


def conv_input_vjp(w, v, input_size, stride=1, padding=0, dilation=1, groups=1) -> np.ndarray:
    w, v = _f64(w), _f64(v)
    n = w.ndim - 2
    input_size = _tup(input_size, n)
    B = v.shape[0]
    g = geometry(B, groups, w.shape[1], w.shape[0] // groups, input_size, w.shape[2:], stride,
                 padding, dilation)
    assert list(v.shape[2:]) == _out_sizes(g), (v.shape, _out_sizes(g))
    gx = np.empty((B, groups * w.shape[1], *input_size), dtype=np.float64)
    lib().oracle_conv_input_vjp(_p(w), _p(v), _p(gx), ctypes.byref(g))
    return gx


def conv_weight_vjp(x, v, kernel_size, dilation=1, padding=0, stride=1, groups=1) -> np.ndarray:
    x, v = _f64(x), _f64(v)
    n = x.ndim - 2
    ks = _tup(kernel_size, n)
    B = x.shape[0]
    g = geometry(B, groups, x.shape[1] // groups, v.shape[1] // groups, x.shape[2:], ks, stride,
                 padding, dilation)
    assert list(v.shape[2:]) == _out_sizes(g), (v.shape, _out_sizes(g))
    gw = np.empty((v.shape[1], x.shape[1] // groups, *ks), dtype=np.float64)
    lib().oracle_conv_weight_vjp(_p(x), _p(v), _p(gw), ctypes.byref(g))
    return gw


def _factor(fn, x, kernel_size, stride, padding, dilation, groups, scale) -> np.ndarray:
    x = _f64(x)
    n = x.ndim - 2
    ks = _tup(kernel_size, n)
    B = x.shape[0]
    g = geometry(B, groups, x.shape[1] // groups, 0, x.shape[2:], ks, stride, padding, dilation)
    A = x.shape[1] // groups * int(np.prod(ks))
    out = np.empty((groups, A, A), dtype=np.float64)
    fn(_p(x), _p(out), ctypes.byref(g), ctypes.c_double(scale))
    return out


def kfc(x, kernel_size, stride=1, padding=0, dilation=1, groups=1, scale=None) -> np.ndarray:
    """scale defaults to 1/B (convNd_kfc.py:105)."""
    if scale is None:
        scale = 1.0 / np.asarray(x).shape[0]
    return _factor(lib().oracle_kfc, x, kernel_size, stride, padding, dilation, groups, scale)


def kfac_reduce(x, kernel_size, stride=1, padding=0, dilation=1, groups=1, scale=None) -> np.ndarray:
    """scale defaults to 1/(B * prod(out)^2) (convNd_kfac_reduce.py:106-108)."""
    if scale is None:
        xs = np.asarray(x).shape
        g = geometry(xs[0], groups, xs[1] // groups, 0, xs[2:], kernel_size, stride, padding, dilation)
        ot = int(np.prod(_out_sizes(g)))
        scale = 1.0 / (xs[0] * ot**2)
    return _factor(lib().oracle_kfac_reduce, x, kernel_size, stride, padding, dilation, groups, scale)


# ---- transpose convolutions ---------------------------------------------
def conv_input_size(out_size, kernel_size, stride, padding, output_padding, dilation) -> int:
    """Input size of the associated convolution (``einconv/utils.py:163-205``), with its check."""
    i = (out_size - 1) * stride - 2 * padding + dilation * (kernel_size - 1) + 1 + output_padding
    if out_size != output_size(i, kernel_size, stride, padding, dilation):
        raise ValueError(f"Output size {out_size} does not match re-computed output size.")
    return i


def _transpose_geometry(x_shape, groups, kernel_size, stride, padding, output_padding, dilation) -> Geom:
    n = len(x_shape) - 2
    ks, ss, ds = _tup(kernel_size, n), _tup(stride, n), _tup(dilation, n)
    ps, ops = _tup(padding, n), _tup(output_padding, n)
    in_sizes = [conv_input_size(x_shape[2 + d], ks[d], ss[d], ps[d], ops[d], ds[d]) for d in range(n)]
    g = geometry(x_shape[0], groups, x_shape[1] // groups, 0, in_sizes, ks, ss, ps, ds)
    assert _out_sizes(g) == list(x_shape[2:]), (_out_sizes(g), x_shape)
    return g


def unfold_transpose(x: np.ndarray, kernel_size, stride=1, padding=0, output_padding=0, dilation=1) -> np.ndarray:
    """Bit-exact input unfolding of a transpose convolution: ``x`` [B, C, *out] -> [B, C*Kt, It]
    (``conv_transposeNd_unfold.py:81-91``)."""
    x = np.ascontiguousarray(x)
    g = _transpose_geometry(x.shape, 1, kernel_size, stride, padding, output_padding, dilation)
    n = g.n_dims
    kt = int(np.prod([g.k[d] for d in range(n)]))
    it = int(np.prod([g.in_[d] for d in range(n)]))
    out = np.empty((x.shape[0], x.shape[1] * kt, it), dtype=x.dtype)
    lib().oracle_unfold_transpose(_p(x), _p(out), ctypes.c_int(x.dtype.itemsize), ctypes.byref(g))
    return out


def kfac_reduce_transpose(x, kernel_size, stride=1, padding=0, output_padding=0, dilation=1, groups=1,
                          scale=None) -> np.ndarray:
    """Input-based KFAC-reduce factor of a transpose convolution, scale defaults to
    1 / (B * prod(in)^2) (``conv_transposeNd_kfac_reduce.py:143-145``)."""
    x = _f64(x)
    g = _transpose_geometry(x.shape, groups, kernel_size, stride, padding, output_padding, dilation)
    n = g.n_dims
    kt = int(np.prod([g.k[d] for d in range(n)]))
    it = int(np.prod([g.in_[d] for d in range(n)]))
    A = x.shape[1] // groups * kt
    if scale is None:
        scale = 1.0 / (x.shape[0] * it**2)
    out = np.empty((groups, A, A), dtype=np.float64)
    lib().oracle_kfac_reduce_transpose(_p(x), _p(out), ctypes.byref(g), ctypes.c_double(scale))
    return out
