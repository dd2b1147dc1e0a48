"""CPU-side checks of the C-ABI library: it loads and exports every symbol include/rbpr.h
declares (no compute without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "rbpr.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rbpr_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_bound_by_ctypes_layer():
    from rbpr import native
    assert sorted(native.SYMBOLS) == _declared()


def test_library_loads_and_exports_every_declared_symbol():
    from rbpr import native
    if not native.LIB_PATH.exists():
        pytest.fail(f"{native.LIB_PATH} missing: run __graft_entry__.build()")
    lib = ctypes.CDLL(str(native.LIB_PATH))
    for name in _declared():
        assert hasattr(lib, name), name
    assert native.load().rbpr_abi_version() == native.ABI_VERSION


def test_hparams_layout_matches_header():
    from rbpr import native
    assert ctypes.sizeof(native.HParams) == 48
    assert native.HParams.lr.offset == 8 and native.HParams.reg_user.offset == 24


def test_create_fails_loudly_without_gpu():
    import torch
    from rbpr import native
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = native.load()
    ctx = ctypes.c_void_p()
    assert lib.rbpr_create(0, ctypes.byref(ctx)) != 0
    from rbpr.engine import Engine
    with pytest.raises(native.NativeError):
        Engine(torch.zeros(4, 8), torch.zeros(4, 8))
