"""CPU-side checks of the C-ABI boundary: the library loads without a GPU and exports every symbol that
include/flexynesis_b200.h declares; argument validation errors surface through fxn_last_error()."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "flexynesis_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fxn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from flexynesis_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(_lib.lib, s), f"libfxn_b200.so does not export {s}"
    assert set(_lib.SYMBOLS) == set(syms), (sorted(set(_lib.SYMBOLS) ^ set(syms)))


def test_version_and_error_reporting():
    from flexynesis_b200 import _lib
    assert _lib.lib.fxn_version() >= 100
    rc = _lib.lib.fxn_gemm(None, None)
    assert rc < 0
    assert b"null" in _lib.lib.fxn_last_error()
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.nterms = 0, 8, 8, 3
    assert _lib.lib.fxn_gemm(ctypes.byref(d), None) < 0
    assert b"positive" in _lib.lib.fxn_last_error()


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import importlib
    from flexynesis_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib._load()
    except ImportError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("missing library must raise")
