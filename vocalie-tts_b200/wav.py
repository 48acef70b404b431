"""PCM_16 WAV container I/O (the wire format on both sides of the hot path: reference
tts_backends/chatterbox_runner.py:152, backend/shared/tts_pipeline.py:409,
backend/shared/audio_edit.py:70).  The sample conversion itself runs on the GPU
(vt_pcm16_encode / vt_pcm16_decode); this module only moves int16 frames."""
from __future__ import annotations

import wave
from pathlib import Path

import numpy as np


def read_pcm16(path) -> tuple[np.ndarray, int]:
    """Return (int16 samples [n] mono, sample_rate).  Multi-channel files are rejected:
    the TTS path is mono."""
    with wave.open(str(path), "rb") as w:
        nch, sw, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if sw != 2:
        raise ValueError(f"{path}: only PCM_16 WAV is supported (sample width {sw})")
    if nch != 1:
        raise ValueError(f"{path}: only mono audio is supported ({nch} channels)")
    return np.frombuffer(raw, dtype="<i2").astype(np.int16, copy=True), int(sr)   # writable (torch.from_numpy)


def write_pcm16(path, samples: np.ndarray, sr: int) -> None:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    q = np.ascontiguousarray(samples, dtype="<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sr))
        w.writeframes(q.tobytes())
