"""Error type at the drop-in boundary.

Everything that goes wrong below ``TTSBackend.synthesize`` surfaces in the reference as
``BackendUnavailableError(RuntimeError)`` (reference tts_backends/base.py:220,
tts_backends/base_runner.py:220-272).  When this package is installed inside a Vocalie-TTS
checkout the reference's own class is re-used so ``except BackendUnavailableError`` in the
pipeline (backend/shared/tts_pipeline.py:14,296-299) catches ours too.
"""
try:  # inside a Vocalie-TTS checkout
    from tts_backends.base import BackendUnavailableError  # type: ignore
except Exception:  # standalone (tests, bench, GPU box)
    class BackendUnavailableError(RuntimeError):
        """Raised when a backend is selected but not available or not wired."""
