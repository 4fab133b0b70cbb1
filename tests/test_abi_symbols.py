"""The C-ABI library loads, exports every symbol include/squava_b200.h declares, and refuses
to compute without a device (no CPU fallback).  No GPU needed."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "squava_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sq_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(sq):
    assert declared_symbols() == sorted(sq.EXPORTS)


def test_library_exports_every_declared_symbol(sq):
    L = sq.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert L.sq_abi_version() == 3


def test_no_cpu_fallback_without_init(sq):
    import numpy as np
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a device is present; covered by the gpu tests")
    sq.shutdown()
    with pytest.raises(sq.SquavaError) as e:
        sq.init()
    assert e.value.status == -4                       # SQ_ERR_NO_DEVICE
    boards = np.zeros(4, sq.BOARD)
    with pytest.raises(sq.SquavaError) as e:
        sq.classify(boards)
    assert e.value.status == -2                       # SQ_ERR_NOT_INIT
    with pytest.raises(sq.SquavaError):
        sq.playouts(boards, np.ones(4, np.int8), 10)
    with pytest.raises(sq.SquavaError):
        sq.alphabeta_root(boards, 1, 4, np.zeros(25, np.int8))
    assert sq.lib().sq_strerror(-2).decode().startswith("library not initialised")


def test_missing_library_fails_loudly(sq, monkeypatch):
    import squava_b200
    monkeypatch.setattr(squava_b200, "_lib", None)
    monkeypatch.setattr(squava_b200, "lib_path", "/nonexistent/libsquava_b200.so")
    with pytest.raises(ImportError):
        squava_b200.lib()
