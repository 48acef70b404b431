"""vocalie-tts_b200 - B200-native HiFT vocoder + post-processing path for Vocalie-TTS.

Only what the hot path needs: ``csrc/`` (CUDA kernels + the C ABI declared in
``include/vocalie_b200.h``), the ctypes binding, and host-side mirrors of the reference
interfaces on this path (``post`` <-> backend/shared/tts_pipeline.py + audio_edit.py,
``backend`` <-> tts_backends/base.py + chatterbox_backend.py).

Import as ``vocalie_tts_b200`` (the hyphenated directory is aliased by the shim package).
"""
from __future__ import annotations

from .errors import BackendUnavailableError
from ._lib import load_library, LIB_PATH, EXPORTED_SYMBOLS

__all__ = ["BackendUnavailableError", "load_library", "LIB_PATH", "EXPORTED_SYMBOLS", "smoke_check"]


def smoke_check() -> None:
    """One small invocation of the hot path on cuda:0 checked against the oracle
    (called by ``__graft_entry__.smoke()``)."""
    from ._smoke import run
    run()
