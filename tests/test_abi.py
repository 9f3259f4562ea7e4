"""The C-ABI library loads and exports every symbol include/brq.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

import breseq_b200 as bq
import helpers


def header_symbols():
    text = open(os.path.join(helpers.ROOT, "include", "brq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(brq_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(built):
    lib = ctypes.CDLL(bq.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libbrq.so does not export %s" % s
    assert sorted(bq.EXPORTS) == syms, "python binding and header disagree"


def test_host_only_context_refuses_compute(built):
    ctx = bq.Context(device=-1)
    with pytest.raises(bq.BrqError):
        ctx.upload()
    with pytest.raises(bq.BrqError):
        ctx.error_count("read_set=1,obs_base,ref_base,quality=42")
    ctx.close()


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(bq, "_lib", None)
    monkeypatch.setattr(bq, "LIB_PATH", "/nonexistent/libbrq.so")
    with pytest.raises(bq.BrqError):
        bq.load_library()


def test_no_device_fails_loudly(built):
    """On a machine without a GPU a device context must raise, never fall back to the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(bq.BrqError):
        bq.Context(device=0)
