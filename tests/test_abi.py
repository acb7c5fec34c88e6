"""The C-ABI library loads without a GPU and exports every symbol include/avhot.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "avhot.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avhot_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_stages():
    names = _declared()
    for want in ("avhot_frontend", "avhot_lstm_forward", "avhot_frame_hidden", "avhot_frame_tail",
                 "avhot_fusion_forward", "avhot_xattn_forward", "avhot_mc_reduce", "avhot_version"):
        assert want in names


def test_library_exports_every_declared_symbol():
    import multimodal_emotion_detection_b200 as pkg
    lib = pkg._lib.load()
    raw = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in avhot.h but not exported"
    assert set(pkg._lib.PROTOTYPES) == set(_declared())
    assert lib.avhot_version() == 100


def test_argument_errors_without_gpu():
    import multimodal_emotion_detection_b200 as pkg
    lib = pkg._lib.load()
    # NULL pointers are rejected before any CUDA call
    assert lib.avhot_frontend(None, None, 1, 48000, None) == pkg._lib.ERR_INVALID
    assert b"frontend" in lib.avhot_last_error()
    assert lib.avhot_mc_reduce(None, None, None, None, None, 1, 1, 8, None) == pkg._lib.ERR_INVALID
    with pytest.raises(ValueError):
        pkg._lib.check(pkg._lib.ERR_INVALID, "x")
    with pytest.raises(RuntimeError):
        pkg._lib.check(pkg._lib.ERR_CUDA, "x")
