"""ctypes binding of ``libeinconv_b200.so`` (the C ABI declared in ``include/einconv_b200.h``).

Only raw device pointers, sizes and a ``cudaStream_t`` cross this boundary; torch is used for
device memory and streams.  There is no CPU path: if the library is missing or a tensor is not
on a CUDA device, the call fails loudly.
"""

import ctypes
import os
from pathlib import Path
from typing import Optional, Sequence

import torch

MAX_DIMS = 8

# einconv_dtype
_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2, torch.float64: 3}
# einconv_math
MATH_AUTO, MATH_FP32, MATH_TF32, MATH_HALF = 0, 1, 2, 3
# einconv_kernels
KERNELS_SPECIALIZED, KERNELS_GENERIC = 0, 1
# einconv_op
OP_UNFOLD, OP_CONV_FWD, OP_CONV_INPUT_VJP, OP_CONV_WEIGHT_VJP, OP_KFC, OP_KFAC_REDUCE, OP_FOLD = range(7)
OP_UNFOLD_TRANSPOSE, OP_KFAC_REDUCE_TRANSPOSE = 7, 8


class Geom(ctypes.Structure):
    """Mirror of ``einconv_geom``."""

    _fields_ = [
        ("n_dims", ctypes.c_int32),
        ("groups", ctypes.c_int32),
        ("batch", ctypes.c_int64),
        ("c_in_g", ctypes.c_int64),
        ("c_out_g", ctypes.c_int64),
        ("in_", ctypes.c_int32 * MAX_DIMS),
        ("k", ctypes.c_int32 * MAX_DIMS),
        ("stride", ctypes.c_int32 * MAX_DIMS),
        ("dilation", ctypes.c_int32 * MAX_DIMS),
        ("pad_left", ctypes.c_int32 * MAX_DIMS),
        ("out", ctypes.c_int32 * MAX_DIMS),
    ]


def make_geom(
    batch: int,
    groups: int,
    c_in_g: int,
    c_out_g: int,
    in_: Sequence[int],
    k: Sequence[int],
    stride: Sequence[int],
    dilation: Sequence[int],
    pad_left: Sequence[int],
    out: Sequence[int],
) -> Geom:
    n = len(in_)
    if not 1 <= n <= MAX_DIMS:
        raise ValueError(f"Number of spatial dimensions must be in [1, {MAX_DIMS}]. Got {n}.")
    g = Geom()
    g.n_dims, g.groups, g.batch, g.c_in_g, g.c_out_g = n, groups, batch, c_in_g, c_out_g
    for d in range(n):
        g.in_[d], g.k[d], g.stride[d] = in_[d], k[d], stride[d]
        g.dilation[d], g.pad_left[d], g.out[d] = dilation[d], pad_left[d], out[d]
    return g


_LIB: Optional[ctypes.CDLL] = None
LIB_PATH = Path(__file__).resolve().parent / "_lib" / "libeinconv_b200.so"

_VOIDP, _I64, _INT, _DBL = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
_GEOMP = ctypes.POINTER(Geom)

# name -> (restype, argtypes); must list every symbol include/einconv_b200.h declares
SIGNATURES = {
    "einconv_abi_version": (_INT, []),
    "einconv_last_error": (ctypes.c_char_p, []),
    "einconv_launch_count": (_I64, []),
    "einconv_selected_kernel": (ctypes.c_char_p, [_INT, _GEOMP, _INT, _INT, _INT]),
    "einconv_workspace_bytes": (_I64, [_INT, _GEOMP, _INT, _INT, _INT]),
    "einconv_unfold": (_INT, [_VOIDP, ctypes.POINTER(_I64), _VOIDP, _INT, _GEOMP, _VOIDP]),
    "einconv_fold": (_INT, [_VOIDP, _VOIDP, _INT, _GEOMP, _VOIDP]),
    "einconv_conv_fwd": (
        _INT,
        [_VOIDP, _VOIDP, _VOIDP, _VOIDP, _INT, _GEOMP, _INT, _INT, _VOIDP, _I64, _VOIDP],
    ),
    "einconv_conv_input_vjp": (
        _INT,
        [_VOIDP, _VOIDP, _VOIDP, _INT, _GEOMP, _INT, _INT, _VOIDP, _I64, _VOIDP],
    ),
    "einconv_conv_weight_vjp": (
        _INT,
        [_VOIDP, _VOIDP, _VOIDP, _INT, _GEOMP, _INT, _INT, _VOIDP, _I64, _VOIDP],
    ),
    "einconv_kfc": (_INT, [_VOIDP, _VOIDP, _INT, _GEOMP, _DBL, _INT, _INT, _VOIDP, _I64, _VOIDP]),
    "einconv_kfac_reduce": (
        _INT,
        [_VOIDP, _VOIDP, _INT, _GEOMP, _DBL, _INT, _INT, _VOIDP, _I64, _VOIDP],
    ),
    "einconv_kfac_reduce_sums": (_INT, [_VOIDP, _VOIDP, _INT, _GEOMP, _VOIDP]),
    "einconv_kfac_reduce_from_sums": (
        _INT,
        [_VOIDP, _INT, _I64, _VOIDP, _INT, _GEOMP, _DBL, _INT, _INT, _VOIDP, _I64, _VOIDP],
    ),
    "einconv_kfac_reduce_from_sums_workspace": (_I64, [_INT, _INT, _GEOMP, _INT, _INT]),
    "einconv_unfold_transpose": (_INT, [_VOIDP, _VOIDP, _INT, _GEOMP, _VOIDP]),
    "einconv_kfac_reduce_transpose": (
        _INT,
        [_VOIDP, _VOIDP, _INT, _GEOMP, _DBL, _INT, _INT, _VOIDP, _I64, _VOIDP],
    ),
}


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _LIB
    if _LIB is None:
        path = Path(os.environ.get("EINCONV_B200_LIB", LIB_PATH))
        if not path.exists():
            raise RuntimeError(
                f"{path} not found: build it with `python -m einconv_b200.build` "
                "(nvcc, sm_100a). einconv_b200 has no fallback without its CUDA library."
            )
        handle = ctypes.CDLL(str(path))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        if handle.einconv_abi_version() != 1:
            raise RuntimeError("libeinconv_b200.so ABI version mismatch; rebuild the library")
        _LIB = handle
    return _LIB


class NativeError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().einconv_last_error().decode(errors="replace")
        if status == -1:
            raise ValueError(f"{what}: {msg}")
        if status == -2:
            raise NotImplementedError(f"{what}: {msg}")
        raise NativeError(f"{what} failed with status {status}: {msg}")


def dtype_code(dtype: torch.dtype) -> int:
    try:
        return _DTYPES[dtype]
    except KeyError:
        raise NotImplementedError(
            f"einconv_b200 kernels support float32, bfloat16, float16 and float64. Got {dtype}."
        ) from None


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors must live on one CUDA device; there is no CPU path."""
    device = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "einconv_b200 evaluates convolution tensor networks with sm_100a CUDA kernels "
                f"only; got a tensor on {t.device}. There is no CPU fallback."
            )
        dev = t.device
        if dev.index is None:  # e.g. a lazy index pattern created with device="cuda": the current device
            dev = torch.device("cuda", torch.cuda.current_device())
        if device is None:
            device = dev
        elif dev != device:
            raise RuntimeError(f"Tensors on different devices: {device} and {dev}.")
    return device


def stream_ptr(device: torch.device) -> int:
    """``cudaStream_t`` of torch's current stream on ``device`` (raw handle: 0.1 us instead of the 2 us of
    building a ``torch.cuda.Stream`` object per call)."""
    index = device.index
    if index is None:
        index = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(index)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def launch_count() -> int:
    return int(lib().einconv_launch_count())
