"""Minimal stand-in for the ``soundfile`` module (absent from the image, no network).

TEST INFRASTRUCTURE.  Lets the reference's file-level functions
(``minimal_post_process``, ``apply_minimal_edit``) run unchanged in the build
container when generating golden vectors.  WAV PCM_16 only, with libsndfile's default
float<->int16 rule (write ``lrintf(x*32767)`` wrapped to 16 bits, read ``q/32768``).
That rule is restated from knowledge of libsndfile, not verified against it here.
"""
from __future__ import annotations

import types
import wave

import numpy as np


def _write(path, data, samplerate, subtype=None, **_kw):
    if subtype not in (None, "PCM_16"):
        raise ValueError(f"sf_stub: unsupported subtype {subtype}")
    a = np.asarray(data)
    if a.ndim == 1:
        nch = 1
    else:
        nch = a.shape[1]
    if a.dtype.kind == "f":
        v = np.rint(a.astype(np.float32) * np.float32(32767.0)).astype(np.int64)
        q = (v & 0xFFFF).astype(np.uint16).view(np.int16)
    elif a.dtype == np.int16:
        q = a
    else:
        raise ValueError("sf_stub: unsupported dtype")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(nch)
        w.setsampwidth(2)
        w.setframerate(int(samplerate))
        w.writeframes(np.ascontiguousarray(q).astype("<i2").tobytes())


def _read(path, dtype="float64", always_2d=False, **_kw):
    with wave.open(str(path), "rb") as w:
        nch, sw, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if sw != 2:
        raise ValueError("sf_stub: PCM_16 only")
    q = np.frombuffer(raw, dtype="<i2")
    if nch > 1 or always_2d:
        q = q.reshape(-1, nch)
    if dtype == "int16":
        return q.copy(), sr
    a = q.astype(np.float64) / 32768.0
    return a.astype(dtype), sr


class _Info:
    def __init__(self, path):
        with wave.open(str(path), "rb") as w:
            self.frames = w.getnframes()
            self.samplerate = w.getframerate()
            self.channels = w.getnchannels()


def make_module() -> types.ModuleType:
    m = types.ModuleType("soundfile")
    m.write = _write
    m.read = _read
    m.info = _Info
    return m
