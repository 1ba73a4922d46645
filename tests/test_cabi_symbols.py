"""CPU: libmsda_b200.so loads and exports every symbol include/msda_b200.h declares; the
Python mirror refuses CPU tensors exactly like the reference (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import dfvod_b200
from dfvod_b200 import _lib
from dfvod_b200 import MultiScaleDeformableAttention as MSDA

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "msda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for name in ("msda_forward", "msda_backward", "msda_abi_version", "msda_error_string"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/msda_b200.h but not exported"


def test_abi_version_and_binding_load():
    assert _lib.load().msda_abi_version() == _lib.ABI_VERSION


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "msda_b200.h")).read()
    assert "at::" not in text and "torch" not in text.lower().replace("pytorch", "")


def _cpu_args():
    shapes = torch.tensor([[2, 2]])
    return (torch.zeros(1, 4, 2, 4), shapes, torch.tensor([0]), torch.zeros(1, 3, 2, 1, 2, 2),
            torch.zeros(1, 3, 2, 1, 2), 64)


def test_cpu_tensors_are_refused_like_the_reference():
    # reference: AT_ERROR("Not implemented on the CPU")  (models/ops/src/ms_deform_attn.h:38,60)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_forward(*_cpu_args())
    a = _cpu_args()
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_backward(*a[:5], torch.zeros(1, 3, 8), 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        dfvod_b200.MSDeformAttnFunction.apply(*a)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmsda_b200.so")
    with pytest.raises(_lib.MSDAError, match="no CPU or PyTorch fallback|not found"):
        _lib.load()
