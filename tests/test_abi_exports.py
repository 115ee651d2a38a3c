"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/hfb200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hfb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hfb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("hfb_dgemm", "hfb_csr_spmm", "hfb_coldot", "hfb_colsum", "hfb_dgemm_batched_small"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from hippyflow_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), "libhfb200.so does not export %s" % s
    assert sorted(_lib.EXPORTED) == declared_symbols()
    assert _lib.lib().hfb_version() >= 100


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call."""
    from hippyflow_b200 import _lib
    L = _lib.lib()
    assert L.hfb_dgemm(7, 4, 4, 4, 1.0, None, 4, None, 4, None, 4, None, 0, 0, None) == -1
    assert L.hfb_dgemm(0, 4, 4, 4, 1.0, None, 4, None, 4, None, 4, None, 0, 0, None) == -1
    assert L.hfb_dgemm(0, 4, 4, 4, 1.0, 16, 3, 32, 4, 64, 4, None, 0, 0, None) == -1   # lda < K
    assert L.hfb_dgemm(0, 4, 4, 4, 1.0, 16, 5, 32, 4, 64, 4, None, 0, 0, None) == -2   # odd lda
    assert L.hfb_dgemm(0, 4, 4, 4, 1.0, 24, 4, 32, 4, 64, 4, None, 0, 0, None) == -2   # misaligned A
    assert L.hfb_csr_spmm(0, 4, None, None, None, None, 4, None, 4, None) == -1
    assert L.hfb_dgemm_workspace_bytes(0, 128, 16, 4096, 4) == 4 * 128 * 16 * 8
    assert L.hfb_dgemm_auto_splits(0, 4096, 266, 263169) >= 2


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (tier rule 3)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hippyflow_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
