"""The C-ABI library loads on a GPU-less host and exports every symbol include/*.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    names = []
    for h in sorted((ROOT / "include").glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names += re.findall(r"\b(vt_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import vocalie_tts_b200 as vt
    return vt.load_library()


def test_header_symbols_are_exported(lib):
    import vocalie_tts_b200 as vt
    declared = _declared_symbols()
    assert declared, "no declarations found in include/*.h"
    raw = ctypes.CDLL(str(vt.LIB_PATH))
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/ but not exported by {vt.LIB_PATH.name}"
    assert sorted(vt.EXPORTED_SYMBOLS) == declared, "ctypes signature table out of sync with the header"


def test_abi_version_and_error_slot(lib):
    assert lib.vt_abi_version() >= 1
    assert isinstance(lib.vt_last_error(), bytes)
    # argument validation happens before any CUDA call, so it works without a GPU
    assert lib.vt_post_workspace_bytes(-1, 0) < 0
    assert lib.vt_post_workspace_bytes(4, 1 << 20) > 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import vocalie_tts_b200 as vt
    from vocalie_tts_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(vt.BackendUnavailableError):
        _lib.load_library()


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import numpy as np
    import vocalie_tts_b200 as vt
    from vocalie_tts_b200 import post
    with pytest.raises(vt.BackendUnavailableError):
        post._find_active_range(np.zeros(8, np.float32), threshold=0.002, min_silence_frames=0)
