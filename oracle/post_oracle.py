"""Post-processing oracle (numpy, CPU).  TEST INFRASTRUCTURE - not a product path.

A restatement, from the behaviour of the reference, of Vocalie-TTS's per-chunk
post-processing.  Every function cites the reference lines it follows
(paths relative to the reference repo root).  Pinned by ``tests/golden/post_*.npz``
which ``oracle/make_golden.py`` produced by running the reference's own functions
(imported unchanged, ``librosa``/``soundfile`` stubbed) in the build container, and by
the reference's four known-answer tests (``tests/test_audio_edges.py:6-27``).

The restatement is written around *index arithmetic* rather than the reference's
array mutations so that it doubles as the specification of the CUDA kernels:
every output sample is ``f32(f32(x * ramp_in) * ramp_out) * f32(scale)`` with the
ramps/scale defined below.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

SILENCE_THRESHOLD = 0.002   # backend/shared/audio_defaults.py:3
SILENCE_MIN_MS = 20         # backend/shared/audio_defaults.py:4
TARGET_SR = 24000           # backend/shared/tts_pipeline.py:26


def ms_to_frames(sr: int, ms) -> int:
    """``int(sr * (int(ms) / 1000.0))`` - tts_pipeline.py:173-174,232,238,246."""
    return max(0, int(sr * (int(ms) / 1000.0)))


# --------------------------------------------------------------------------- ramps
def ramp_in(n: int) -> np.ndarray:
    """Fade-in ramp of ``_fade_in`` (tts_pipeline.py:140-148).

    ``np.linspace(0, 1, n, endpoint=True, dtype=float32)``: element i is
    ``float32(float64(i) * (1/(n-1)))`` with the last element forced to 1.0
    (numpy's linspace: ``y = arange(n) * step + start`` in float64, ``y[-1] = stop``,
    then cast).  ``n == 1`` gives ``[0.0]``.
    """
    if n <= 0:
        return np.zeros(0, np.float32)
    if n == 1:
        return np.zeros(1, np.float32)
    step = 1.0 / (n - 1)
    y = np.arange(n, dtype=np.float64) * step
    y[-1] = 1.0
    return y.astype(np.float32)


def ramp_out(n: int) -> np.ndarray:
    """Fade-out ramp of ``_fade_out`` (tts_pipeline.py:151-159): linspace(1, 0, n)."""
    if n <= 0:
        return np.zeros(0, np.float32)
    if n == 1:
        return np.ones(1, np.float32)
    step = -1.0 / (n - 1)
    y = np.arange(n, dtype=np.float64) * step + 1.0
    y[-1] = 0.0
    return y.astype(np.float32)


def fade_in(audio: np.ndarray, fade_frames: int) -> np.ndarray:
    """In-place head fade - tts_pipeline.py:140-148."""
    if audio.size == 0:
        return audio
    f = max(0, min(int(fade_frames), len(audio)))
    if f == 0:
        return audio
    audio[:f] *= ramp_in(f)
    return audio


def fade_out(audio: np.ndarray, fade_frames: int) -> np.ndarray:
    """In-place tail fade - tts_pipeline.py:151-159."""
    if audio.size == 0:
        return audio
    f = max(0, min(int(fade_frames), len(audio)))
    if f == 0:
        return audio
    audio[len(audio) - f:] *= ramp_out(f)
    return audio


# --------------------------------------------------------------------------- trim
def find_active_range(mono: np.ndarray, *, threshold: float, min_silence_frames: int):
    """First/last sample with ``|x| > threshold`` - tts_pipeline.py:192-209.

    The compare is done in float32 (NEP-50: the Python float is cast to the array
    dtype), so the effective threshold is ``float32(threshold)``.
    """
    n = int(mono.size)
    if n == 0:
        return 0, 0
    thr = np.float32(threshold) if mono.dtype == np.float32 else float(threshold)
    active = np.flatnonzero(np.abs(mono) > thr)
    if active.size == 0:
        return 0, n
    start = int(active[0])
    end = int(active[-1]) + 1
    if start < min_silence_frames:
        start = 0
    if n - end < min_silence_frames:
        end = n
    return start, end


def snap_zero_crossing(audio: np.ndarray, idx: int, *, radius_samples: int) -> int:
    """Nearest zero crossing within +-radius - tts_pipeline.py:114-137.

    Vectorised equivalent of the reference's ascending Python loop with a strict
    ``dist < best_dist`` update: among crossing positions i in
    ``[max(idx-r,1), min(idx+r,N-1)]`` pick minimal ``|i-idx|``, ties -> lower i;
    none -> idx (after clamping idx to ``[0, N-1]``).
    """
    n = int(audio.size)
    if n == 0:
        return idx
    idx = max(min(int(idx), n - 1), 0)
    lo = max(idx - radius_samples, 1)
    hi = min(idx + radius_samples, n - 1)
    if hi < lo:
        return idx
    prev = audio[lo - 1:hi].astype(np.float64)
    cur = audio[lo:hi + 1].astype(np.float64)
    cross = (prev == 0.0) | (cur == 0.0) | ((prev < 0.0) & (cur >= 0.0)) | ((prev > 0.0) & (cur <= 0.0))
    cand = np.flatnonzero(cross)
    if cand.size == 0:
        return idx
    pos = cand + lo
    dist = np.abs(pos - idx)
    # best_dist starts at radius+1, every candidate is within radius -> always accepted
    best = int(np.argmin(dist))  # argmin returns the first (lowest index) minimum
    return int(pos[best])


def trim_range_snapped(mono: np.ndarray, sr: int, *, silence_threshold=SILENCE_THRESHOLD,
                       silence_min_ms=SILENCE_MIN_MS, zero_cross_radius_ms=10):
    """(start, end) exactly as ``minimal_post_process`` computes them - tts_pipeline.py:232-244."""
    n = int(mono.size)
    min_sil = int(sr * (int(silence_min_ms) / 1000.0))
    start, end = find_active_range(mono, threshold=float(silence_threshold), min_silence_frames=min_sil)
    radius = int(sr * (int(zero_cross_radius_ms) / 1000.0))
    if n:
        start = snap_zero_crossing(mono, start, radius_samples=radius)
        end = snap_zero_crossing(mono, max(end - 1, start), radius_samples=radius) + 1
    if end <= start:
        start, end = 0, n
    return start, end


# --------------------------------------------------------------------------- normalise
def peak_scale(peak: float, target_db: float):
    """``scale = 10**(dB/20) / peak`` in float64 - tts_pipeline.py:254-259, audio_edit.py:58-66."""
    target_peak = float(10 ** (float(target_db) / 20.0))
    if peak > 0.0 and target_peak > 0.0:
        return target_peak / peak
    return 1.0


def minimal_post_process_array(audio: np.ndarray, sr: int, *, zero_cross_radius_ms=10, fade_ms=10,
                               silence_threshold=SILENCE_THRESHOLD, silence_min_ms=SILENCE_MIN_MS,
                               normalize_peak_db=-1.0):
    """Array-level body of ``minimal_post_process`` (tts_pipeline.py:229-274), mono.

    Returns ``(processed float32, meta)`` where meta has the reference's keys.
    """
    audio = np.asarray(audio, dtype=np.float32)
    start, end = trim_range_snapped(audio, sr, silence_threshold=silence_threshold,
                                    silence_min_ms=silence_min_ms, zero_cross_radius_ms=zero_cross_radius_ms)
    trimmed = audio[start:end].copy()
    f = int(sr * (int(fade_ms) / 1000.0))
    trimmed = fade_in(trimmed, f)
    trimmed = fade_out(trimmed, f)
    peak_before = float(np.max(np.abs(trimmed))) if trimmed.size else 0.0
    scale = 1.0
    target_peak = float(10 ** (float(normalize_peak_db) / 20.0))
    if peak_before > 0.0 and target_peak > 0.0:
        scale = target_peak / peak_before
        trimmed = trimmed * scale  # float32 array * python float -> float32 (x * f32(scale))
    meta = {
        "trim": {"start_sample": int(start), "end_sample": int(end)},
        "fade_ms": int(fade_ms),
        "zero_cross_radius_ms": int(zero_cross_radius_ms),
        "silence_threshold": float(silence_threshold),
        "silence_min_ms": int(silence_min_ms),
        "normalize_peak_db": float(normalize_peak_db),
        "normalize_scale": float(scale),
        "peak_before": float(peak_before),
    }
    return trimmed, meta


def apply_minimal_edit_array(audio: np.ndarray, sr: int, *, trim_enabled: bool, normalize_enabled: bool,
                             target_dbfs: float, silence_threshold=SILENCE_THRESHOLD,
                             silence_min_ms=SILENCE_MIN_MS):
    """Array-level body of ``apply_minimal_edit`` (audio_edit.py:41-79), mono.

    Returns ``(clipped float32 ready for PCM_16, result dict)``.
    """
    audio = np.asarray(audio, dtype=np.float32)
    trimmed = False
    if trim_enabled:
        min_sil = int(sr * (int(silence_min_ms) / 1000.0))
        s, e = find_active_range(audio, threshold=float(silence_threshold), min_silence_frames=min_sil)
        if 0 <= s < e <= len(audio):
            audio = audio[s:e]
            trimmed = True
    normalized = False
    peak_before = float(np.max(np.abs(audio))) if audio.size else 0.0
    target_peak = 10 ** (float(target_dbfs) / 20.0)
    gain = 1.0
    if normalize_enabled and peak_before > 0.0 and target_peak > 0.0:
        gain = target_peak / peak_before
        audio = audio * gain
        normalized = True
    audio = np.clip(audio, -1.0, 1.0)
    return audio, {
        "trimmed": trimmed,
        "normalized": normalized,
        "target_dbfs": float(target_dbfs),
        "peak_before": peak_before,
        "peak_after": float(np.max(np.abs(audio))) if audio.size else 0.0,
        "gain": gain,
    }


# --------------------------------------------------------------------------- stitch
def apply_inter_chunk_gap(chunks: Sequence[np.ndarray], *, sr: int, gap_ms: int, fade_ms: int = 10) -> np.ndarray:
    """Gap-padded concatenation with edge fades - tts_pipeline.py:162-189.

    Chunk i gets fade-out iff i < last and fade-in iff i > 0, *out before in*;
    ``zeros(gap)`` follows every chunk but the last; ``gap_ms <= 0`` or a single
    chunk is a plain concatenate without fades.
    """
    if not chunks:
        return np.zeros(0, np.float32)
    if gap_ms <= 0 or len(chunks) == 1:
        return np.concatenate(chunks)
    gap = ms_to_frames(sr, gap_ms)
    fade = ms_to_frames(sr, fade_ms)
    last = len(chunks) - 1
    total = sum(int(np.asarray(c).size) for c in chunks) + last * gap
    out = np.zeros(total, np.float32)
    pos = 0
    for i, c in enumerate(chunks):
        a = np.array(c, dtype=np.float32, copy=True)
        if fade > 0:
            if i < last:
                fade_out(a, fade)
            if i > 0:
                fade_in(a, fade)
        out[pos:pos + a.size] = a
        pos += a.size
        if i < last:
            pos += gap
    return out


def stitched_length(lengths: Sequence[int], *, sr: int, gap_ms: int) -> int:
    """Sample count of ``apply_inter_chunk_gap`` - the bit-exact length contract."""
    lengths = [int(v) for v in lengths]
    if not lengths:
        return 0
    if gap_ms <= 0 or len(lengths) == 1:
        return int(sum(lengths))
    return int(sum(lengths)) + (len(lengths) - 1) * ms_to_frames(sr, gap_ms)


# --------------------------------------------------------------------------- PCM_16 wire format
def pcm16_encode(x: np.ndarray) -> np.ndarray:
    """float32 -> int16 as libsndfile writes PCM_16 with its defaults (normalised
    floats, clipping off): ``lrintf(x * 32767)`` truncated to 16 bits.  This is the
    wire format at tts_backends/chatterbox_runner.py:152, tts_pipeline.py:409 and
    audio_edit.py:70.  (From knowledge of libsndfile ``pcm.c``; libsndfile itself is
    absent from the build container, so this rule is *not* pinned by a run.)
    """
    x = np.asarray(x, dtype=np.float32)
    v = np.rint(x * np.float32(32767.0)).astype(np.int64)
    return (v & 0xFFFF).astype(np.uint16).view(np.int16)


def pcm16_decode(q: np.ndarray) -> np.ndarray:
    """int16 -> float32 as ``sf.read(dtype='float32')``: ``q / 32768``
    (tts_backends/base_runner.py:323)."""
    return (np.asarray(q, dtype=np.int16).astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def rms(x: np.ndarray) -> float:
    """``sqrt(mean(x.astype(f64)**2))`` - tts_backends/cosyvoice_backend.py:103 (helper only)."""
    x = np.asarray(x, dtype=np.float64)
    return float(math.sqrt(np.mean(x * x))) if x.size else 0.0
