"""CPU: the C-ABI shared library loads and exports every symbol ``include/einconv_b200.h`` declares.

No compute call is made (there is no GPU here); the host-only entry points (version, workspace
sizing, kernel selection, error text) are exercised.
"""

import ctypes
import re
from pathlib import Path

import pytest

from einconv_b200 import _native as nat

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "einconv_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(einconv_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_six_networks():
    names = declared_symbols()
    for required in ("einconv_unfold", "einconv_conv_fwd", "einconv_conv_input_vjp",
                     "einconv_conv_weight_vjp", "einconv_kfc", "einconv_kfac_reduce",
                     "einconv_workspace_bytes", "einconv_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol():
    handle = ctypes.CDLL(str(nat.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} is declared in the header but not exported"


def test_binding_covers_every_declared_symbol():
    assert sorted(nat.SIGNATURES) == declared_symbols()


def test_host_only_entry_points():
    lib = nat.lib()
    assert lib.einconv_abi_version() == 1
    g = nat.make_geom(4, 1, 8, 8, (16, 16), (3, 3), (1, 1), (1, 1), (1, 1), (16, 16))
    assert lib.einconv_workspace_bytes(nat.OP_UNFOLD, g, 0, 0, 0) == 0
    assert lib.einconv_workspace_bytes(nat.OP_CONV_WEIGHT_VJP, g, 0, nat.MATH_FP32, 0) > 0
    assert lib.einconv_selected_kernel(nat.OP_UNFOLD, g, 0, 0, 0) == b"unfold_plane"
    dw = nat.make_geom(4, 8, 1, 1, (16, 16), (3, 3), (1, 1), (1, 1), (1, 1), (16, 16))
    assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, dw, 0, 0, nat.KERNELS_SPECIALIZED) == b"depthwise_stencil"
    assert lib.einconv_selected_kernel(nat.OP_CONV_FWD, dw, 0, 0, nat.KERNELS_GENERIC) == b"ffma_gather_gemm"


def test_invalid_geometry_reports_an_error():
    lib = nat.lib()
    bad = nat.make_geom(4, 1, 8, 8, (16,), (3,), (0,), (1,), (1,), (16,))  # stride 0
    assert lib.einconv_workspace_bytes(nat.OP_CONV_FWD, bad, 0, 0, 0) < 0
    assert b"invalid geometry" in lib.einconv_last_error()
    with pytest.raises(ValueError):
        nat.check(-1, "probe")
