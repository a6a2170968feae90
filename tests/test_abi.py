"""CPU: the C-ABI library builds, loads, and exports every symbol that
include/srcv_b200.h declares; host-side argument validation works without a GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

from simplerecon_b200 import _native

HEADER = Path(__file__).resolve().parents[1] / "include" / "srcv_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(srcv_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for must in ("srcv_dot_forward_f32", "srcv_mlp_forward_f32", "srcv_dot_workspace_bytes",
                 "srcv_mlp_workspace_bytes", "srcv_abi_version", "srcv_check_device"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    raw = C.CDLL(str(_native.LIB_PATH))
    for name in declared_functions():
        assert hasattr(raw, name), f"{name} declared in srcv_b200.h but not exported"
    # and the binding table covers the header
    assert sorted(_native.SYMBOLS) == declared_functions()


def test_abi_version_and_status_strings(built_lib):
    assert built_lib.srcv_abi_version() == 2
    assert built_lib.srcv_status_string(0) == b"ok"
    assert built_lib.srcv_status_string(2) == b"bad shape"


def test_argument_validation_without_gpu(built_lib):
    lib = built_lib
    bad = _native.Shape(0, 7, 16, 120, 160, 64)
    assert lib.srcv_dot_workspace_bytes(C.byref(bad)) == 0
    good = _native.Shape(4, 7, 16, 120, 160, 64)
    n = lib.srcv_dot_workspace_bytes(C.byref(good))
    # channel-last copy of the sources dominates: B*K*C*H*W floats
    assert n >= 4 * 7 * 16 * 120 * 160 * 4
    cams, pl = _native.Cameras(), _native.Planes()
    st = lib.srcv_dot_forward_f32(C.byref(good), None, None, C.byref(cams), C.byref(pl), None, None,
                                  None, 0, None)
    assert st == 1 and b"NULL" in lib.srcv_last_error()          # SRCV_ERR_NULL
    st = lib.srcv_dot_forward_f32(C.byref(bad), None, None, C.byref(cams), C.byref(pl), None, None,
                                  None, 0, None)
    assert st == 2                                                # SRCV_ERR_SHAPE
    with pytest.raises(_native.SrcvError):
        _native.check(st)
    w = _native.MlpWeights(None, None, None, None, None, None, 128, 128)
    assert lib.srcv_mlp_workspace_bytes(C.byref(good), C.byref(w)) > 0
    w_bad = _native.MlpWeights(None, None, None, None, None, None, 4096, 128)
    assert lib.srcv_mlp_workspace_bytes(C.byref(good), C.byref(w_bad)) == 0
    assert lib.srcv_set_variant(99) == 4                          # SRCV_ERR_UNSUPPORTED
    assert lib.srcv_set_variant(0) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setenv("SRCV_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_native.NativeLibraryError):
        _native.load()
