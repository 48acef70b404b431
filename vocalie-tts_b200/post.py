"""Host-side mirror of the reference's post-processing interface, running on the GPU.

Same names, argument meaning and error behaviour as the reference functions it replaces
(backend/shared/tts_pipeline.py:114-274, backend/shared/audio_edit.py:16-79); the arithmetic
runs in ``csrc/vt_post.cu`` through the C ABI.  There is no CPU fallback: without the CUDA
extension or a CUDA device every call raises ``BackendUnavailableError``.

Two layers:
  * device layer  - ``post_process_device`` & friends: torch CUDA tensors in/out, no host sync
    except the optional read-back of the per-segment results.  Used by the pipeline/bench.
  * reference-facing layer - ``_find_active_range``, ``_snap_zero_crossing``, ``_fade_in``,
    ``_fade_out``, ``_apply_inter_chunk_gap``, ``minimal_post_process``, ``apply_minimal_edit``:
    numpy arrays / WAV paths in and out, exactly like the reference.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import PostParams, check
from .errors import BackendUnavailableError
from . import wav as _wav

SILENCE_THRESHOLD = 0.002   # reference backend/shared/audio_defaults.py:3
SILENCE_MIN_MS = 20         # reference backend/shared/audio_defaults.py:4
TARGET_SR = 24000           # reference backend/shared/tts_pipeline.py:26


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise BackendUnavailableError("no CUDA device: the B200 path has no CPU fallback")
    return torch


def _ms_to_frames(sr: int, ms) -> int:
    return max(0, int(sr * (int(ms) / 1000.0)))   # tts_pipeline.py:173-174,232,238,246


def _ptr(t) -> int:
    return 0 if t is None else int(t.data_ptr())


def _stream(torch) -> int:
    return int(torch.cuda.current_stream().cuda_stream)


@dataclass
class PostResult:
    out: Any                 # torch CUDA tensor (float32 or int16), capacity >= total
    total: Optional[int]     # total output samples (None when not read back)
    results: Optional[np.ndarray]  # [n_seg, 8] float64: start,end,peak,scale,dst,len,peak_used,-
    results_dev: Any = None  # the same table and the total as CUDA tensors (always set: callers that defer the
    total_dev: Any = None    # read-back copy them on their own stream)


class _Workspace:
    """Grow-only device scratch of the post calls, one buffer PER THREAD AND DEVICE.

    ``vt_post_analyze`` leaves its plan in the workspace and ``vt_post_write`` reads it back, and ctypes drops
    the GIL during both calls; the reference runs up to ``MAX_CONCURRENT_JOBS = 2`` job threads
    (backend/config.py:11), so a process-wide buffer would let thread B's analyse overwrite thread A's plan
    between A's two calls.  Thread-local buffers make every analyse -> write pair private without serialising
    the threads (kernels stay stream-ordered on the caller's stream)."""
    _tls = threading.local()

    @classmethod
    def get(cls, torch, nbytes: int, device=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        bufs = getattr(cls._tls, "bufs", None)
        if bufs is None:
            bufs = cls._tls.bufs = {}
        buf = bufs.get(device)
        if buf is None or buf.numel() < nbytes:
            buf = bufs[device] = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        return buf


def last_launch_count() -> int:
    """Kernels launched by the calling thread's last post call (vt_last_launch_count)."""
    return int(_lib.load_library().vt_last_launch_count())


def make_params(*, sr=TARGET_SR, trim=0, silence_threshold=SILENCE_THRESHOLD, min_silence_frames=0,
                snap_radius=-1, fade_in_frames=0, fade_out_frames=0, stitch=0, gap_frames=0,
                normalize=0, clip=0, target_peak=1.0, concat=1, out_pcm16=0, stitch_head=1, stitch_tail=1) -> PostParams:
    return PostParams(int(sr), int(trim), float(silence_threshold), int(min_silence_frames),
                      int(snap_radius), int(fade_in_frames), int(fade_out_frames), int(stitch),
                      int(gap_frames), int(normalize), int(clip), float(target_peak), int(concat),
                      int(out_pcm16), int(stitch_head), int(stitch_tail))


def _seg_arrays(torch, seg_off, device=None):
    seg_np = np.ascontiguousarray(seg_off, dtype=np.int64)
    if seg_np.ndim != 1 or seg_np.size < 1 or np.any(np.diff(seg_np) < 0) or seg_np[0] != 0:
        raise ValueError("seg_off must be a non-decreasing int64 array starting at 0")
    n_seg = seg_np.size - 1
    n_samples = int(seg_np[-1])
    max_len = int(np.max(np.diff(seg_np))) if n_seg else 0
    seg_dev = torch.from_numpy(seg_np).to(device if device is not None else "cuda")
    return seg_np, seg_dev, n_seg, n_samples, max_len


def post_process_device(audio, seg_off, params: PostParams, *, out=None, range_override=None,
                        peak_override=None, read_back=True) -> PostResult:
    """Run analyse + write on device tensors.  ``audio``: 1-D float32 CUDA tensor holding all
    segments; ``seg_off``: int64[n_seg+1] (host).  Output capacity is the worst case
    ``n_samples + (n_seg-1)*gap``.  Everything is allocated and launched on ``audio``'s device and on
    the calling thread's current stream for that device (job threads default to device 0 otherwise)."""
    torch = _torch()
    lib = _lib.load_library()
    if audio.dtype != torch.float32 or audio.dim() != 1 or not audio.is_cuda:
        raise ValueError("audio must be a 1-D float32 CUDA tensor")
    dev = audio.device
    with torch.cuda.device(dev):
        audio = audio.contiguous()
        if audio.data_ptr() % 16:
            audio = audio.clone()
        seg_np, seg_dev, n_seg, n_samples, max_len = _seg_arrays(torch, seg_off, dev)
        if n_samples > audio.numel():
            raise ValueError("seg_off exceeds the audio buffer")
        cap = n_samples + n_seg * max(int(params.gap_frames), 0) if params.concat else n_samples
        odt = torch.int16 if params.out_pcm16 else torch.float32
        if out is None:
            out = torch.empty(max(cap, 4), dtype=odt, device=dev) if params.concat else \
                torch.zeros(max(cap, 4), dtype=odt, device=dev)
        elif out.dtype != odt or out.numel() < cap or not out.is_cuda or out.device != dev:
            raise ValueError("out tensor has the wrong dtype/size/device")
        for name, t in (("range_override", range_override), ("peak_override", peak_override)):
            if t is not None and (not t.is_cuda or t.device != dev):
                raise ValueError(f"{name} must live on the audio's device")
        ws_bytes = int(lib.vt_post_workspace_bytes(n_seg, n_samples))
        ws = _Workspace.get(torch, ws_bytes, dev)
        res = torch.empty((max(n_seg, 1), _lib.POST_RESULT_STRIDE), dtype=torch.float64, device=dev)
        tot = torch.zeros(1, dtype=torch.int64, device=dev)
        st = _stream(torch)
        check(lib.vt_post_analyze(_ptr(audio), _ptr(seg_dev), n_seg, n_samples, max_len, C.byref(params),
                                  _ptr(range_override), _ptr(ws), ws.numel(), st), "vt_post_analyze")
        check(lib.vt_post_write(_ptr(audio), _ptr(seg_dev), n_seg, n_samples, max_len, C.byref(params),
                                _ptr(peak_override), _ptr(out), out.numel(), _ptr(res), _ptr(tot),
                                _ptr(ws), ws.numel(), st), "vt_post_write")
        if not read_back:
            return PostResult(out, None, None, res, tot)
        res_h = res.cpu().numpy()[:n_seg]
        return PostResult(out, int(tot.item()), res_h, res, tot)


# ------------------------------------------------------------------------- WAV file images on the device
WAV_HEADER_BYTES = 44
_WAV_SAMPLES_AT = 48      # header at bytes [4, 48): the samples then start 16-byte aligned for the vectorised writer


class WavImage:
    """A mono PCM_16 WAV file assembled in device memory: ``samples`` is the int16 view the post writer fills
    (``out=``), ``finish`` writes the RIFF header in front of it on the device and ``to_file`` moves the finished image
    to the host in one copy - no float32 bounce, no host-side header (reference wire format:
    tts_backends/chatterbox_runner.py:152, backend/shared/tts_pipeline.py:409, backend/shared/audio_edit.py:70)."""

    def __init__(self, capacity_samples: int, device=None):
        torch = _torch()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.capacity = max(int(capacity_samples), 4)
        self.buf = torch.empty(_WAV_SAMPLES_AT + 2 * self.capacity + 16, dtype=torch.uint8, device=self.device)
        self.samples = self.buf[_WAV_SAMPLES_AT:_WAV_SAMPLES_AT + 2 * self.capacity].view(torch.int16)
        self._host = None

    def finish(self, sr: int, *, total_dev=None, total: Optional[int] = None) -> None:
        """Write the header for ``total`` samples (host int) or ``total_dev`` (device int64[1], e.g. PostResult.total_dev)."""
        torch = _torch()
        with torch.cuda.device(self.device):
            check(_lib.load_library().vt_wav_pcm16_header(int(self.buf.data_ptr()) + 4, int(sr), _ptr(total_dev),
                                                          int(total or 0), _stream(torch)), "vt_wav_pcm16_header")

    def to_host(self, total: int):
        """Pinned host bytes of the finished file (header + ``total`` samples); synchronises the current stream."""
        torch = _torch()
        n = WAV_HEADER_BYTES + 2 * int(total)
        if self._host is None or self._host.numel() < n:
            self._host = torch.empty(max(n, WAV_HEADER_BYTES + 2 * self.capacity), dtype=torch.uint8).pin_memory()
        with torch.cuda.device(self.device):
            self._host[:n].copy_(self.buf[4:4 + n], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return self._host[:n]

    def to_file(self, path, total: int) -> int:
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        data = self.to_host(total).numpy()
        with open(path, "wb") as f:
            f.write(memoryview(data))
        return int(data.size)


def stats_device(audio, seg_off, *, threshold: float = SILENCE_THRESHOLD):
    """One read pass over ``audio`` (float32 CUDA, segments ``seg_off``): returns device tensors
    ``first_last`` int64[n_seg, 2] (first / last sample with |x| > threshold, -1 / -1 if none) and ``peak``
    float32[n_seg].  No host synchronisation."""
    torch = _torch()
    lib = _lib.load_library()
    if audio.dtype != torch.float32 or audio.dim() != 1 or not audio.is_cuda:
        raise ValueError("audio must be a 1-D float32 CUDA tensor")
    dev = audio.device
    with torch.cuda.device(dev):
        seg_np, seg_dev, n_seg, n_samples, max_len = _seg_arrays(torch, seg_off, dev)
        fl = torch.full((max(n_seg, 1), 2), -1, dtype=torch.int64, device=dev)
        pk = torch.zeros(max(n_seg, 1), dtype=torch.float32, device=dev)
        ws = _Workspace.get(torch, int(lib.vt_post_workspace_bytes(n_seg, n_samples)), dev)
        check(lib.vt_post_stats(_ptr(audio), _ptr(seg_dev), n_seg, n_samples, max_len, float(threshold), _ptr(fl),
                                _ptr(pk), _ptr(ws), ws.numel(), _stream(torch)), "vt_post_stats")
    return fl[:n_seg], pk[:n_seg]


# ------------------------------------------------------------------------- reference-facing layer
def _as_f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))


def _find_active_range(mono: np.ndarray, *, threshold: float, min_silence_frames: int) -> tuple[int, int]:
    """GPU ``_find_active_range`` (reference tts_pipeline.py:192-209)."""
    torch = _torch()
    lib = _lib.load_library()
    x = _as_f32(mono)
    if x.size == 0:
        return 0, 0
    xd = torch.from_numpy(x).cuda()
    seg_np, seg_dev, n_seg, n_samples, max_len = _seg_arrays(torch, [0, x.size])
    ws = _Workspace.get(torch, int(lib.vt_post_workspace_bytes(1, n_samples)))
    rng = torch.empty(2, dtype=torch.int64, device="cuda")
    check(lib.vt_find_active_range(_ptr(xd), _ptr(seg_dev), 1, n_samples, max_len, float(threshold),
                                   int(min_silence_frames), _ptr(rng), _ptr(ws), ws.numel(), _stream(torch)),
          "vt_find_active_range")
    s, e = rng.cpu().tolist()
    return int(s), int(e)


def _snap_zero_crossing(audio: np.ndarray, idx: int, *, radius_samples: int) -> int:
    """GPU ``_snap_zero_crossing`` (reference tts_pipeline.py:114-137)."""
    torch = _torch()
    lib = _lib.load_library()
    x = _as_f32(audio)
    if x.size == 0:
        return idx
    xd = torch.from_numpy(x).cuda()
    seg_np, seg_dev, *_ = _seg_arrays(torch, [0, x.size])
    i_in = torch.tensor([int(idx)], dtype=torch.int64, device="cuda")
    i_out = torch.empty(1, dtype=torch.int64, device="cuda")
    check(lib.vt_snap_zero_crossing(_ptr(xd), _ptr(seg_dev), 1, _ptr(i_in), int(radius_samples),
                                    _ptr(i_out), _stream(torch)), "vt_snap_zero_crossing")
    return int(i_out.item())


def _fade(audio: np.ndarray, fade_frames: int, which: str) -> np.ndarray:
    if audio.size == 0:
        return audio
    f = max(0, min(int(fade_frames), len(audio)))
    if f == 0:
        return audio
    torch = _torch()
    x = _as_f32(audio)
    prm = make_params(fade_in_frames=f if which == "in" else 0, fade_out_frames=f if which == "out" else 0)
    r = post_process_device(torch.from_numpy(x).cuda(), [0, x.size], prm, read_back=False)
    audio[...] = r.out[:x.size].cpu().numpy().reshape(audio.shape)   # in place, like the reference
    return audio


def _fade_in(audio: np.ndarray, fade_frames: int) -> np.ndarray:
    """GPU ``_fade_in`` (reference tts_pipeline.py:140-148); mutates and returns ``audio``."""
    return _fade(audio, fade_frames, "in")


def _fade_out(audio: np.ndarray, fade_frames: int) -> np.ndarray:
    """GPU ``_fade_out`` (reference tts_pipeline.py:151-159); mutates and returns ``audio``."""
    return _fade(audio, fade_frames, "out")


def stitch_params(n_chunks: int, *, sr: int, gap_ms: int, fade_ms: int = 10, **kw) -> PostParams:
    """Parameter block with ``_apply_inter_chunk_gap`` semantics (tts_pipeline.py:162-189):
    ``gap_ms <= 0`` or a single chunk is a plain concatenate without fades."""
    if gap_ms <= 0 or n_chunks <= 1:
        return make_params(sr=sr, stitch=1, gap_frames=0, concat=1, **kw)
    fade = _ms_to_frames(sr, fade_ms)
    return make_params(sr=sr, stitch=1, gap_frames=_ms_to_frames(sr, gap_ms), fade_in_frames=fade,
                       fade_out_frames=fade, concat=1, **kw)


def _apply_inter_chunk_gap(audio_chunks: Sequence[np.ndarray], *, sr: int, gap_ms: int, fade_ms: int = 10) -> np.ndarray:
    """GPU ``_apply_inter_chunk_gap`` (reference tts_pipeline.py:162-189)."""
    if not audio_chunks:
        return np.zeros(0, dtype=np.float32)
    torch = _torch()
    chunks = [_as_f32(c) for c in audio_chunks]
    lens = [c.size for c in chunks]
    seg_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
    if flat.size == 0:
        prm0 = stitch_params(len(chunks), sr=sr, gap_ms=gap_ms, fade_ms=fade_ms)
        return np.zeros(max(len(chunks) - 1, 0) * int(prm0.gap_frames), dtype=np.float32)
    prm = stitch_params(len(chunks), sr=sr, gap_ms=gap_ms, fade_ms=fade_ms)
    r = post_process_device(torch.from_numpy(flat).cuda(), seg_off, prm)
    return r.out[:r.total].cpu().numpy()


def minimal_post_process(raw_path, processed_path, *, zero_cross_radius_ms: int = 10, fade_ms: int = 10,
                         silence_threshold: float = SILENCE_THRESHOLD, silence_min_ms: int = SILENCE_MIN_MS,
                         normalize_peak_db: float = -1.0) -> dict[str, Any]:
    """GPU ``minimal_post_process`` (reference tts_pipeline.py:212-274): trim + zero-cross snap +
    fades + peak normalise, PCM_16 WAV in / PCM_16 WAV out, same meta dict."""
    raw_path = Path(raw_path)
    processed_path = Path(processed_path)
    if raw_path.resolve() == processed_path.resolve():
        raise ValueError("Le traitement doit écrire dans un fichier différent du raw.")
    torch = _torch()
    lib = _lib.load_library()
    q, sr = _wav.read_pcm16(raw_path)
    n = q.size
    fade = int(sr * (int(fade_ms) / 1000.0))
    prm = make_params(sr=sr, trim=1, silence_threshold=float(silence_threshold),
                      min_silence_frames=int(sr * (int(silence_min_ms) / 1000.0)),
                      snap_radius=int(sr * (int(zero_cross_radius_ms) / 1000.0)),
                      fade_in_frames=fade, fade_out_frames=fade, normalize=1,
                      target_peak=float(10 ** (float(normalize_peak_db) / 20.0)), out_pcm16=1)
    start = end = 0
    scale, peak = 1.0, 0.0
    out_q = np.zeros(0, np.int16)
    if n:
        qd = torch.from_numpy(q).cuda()
        x = torch.empty(n + 4, dtype=torch.float32, device="cuda")
        check(lib.vt_pcm16_decode(_ptr(qd), _ptr(x), n, _stream(torch)), "vt_pcm16_decode")
        r = post_process_device(x, [0, n], prm)
        start, end, peak, scale = int(r.results[0, 0]), int(r.results[0, 1]), float(r.results[0, 2]), float(r.results[0, 3])
        out_q = r.out[:r.total].cpu().numpy()
    _wav.write_pcm16(processed_path, out_q, sr)
    return {
        "trim": {"start_sample": int(start), "end_sample": int(end)},
        "fade_ms": int(fade_ms),
        "zero_cross_radius_ms": int(zero_cross_radius_ms),
        "silence_threshold": float(silence_threshold),
        "silence_min_ms": int(silence_min_ms),
        "normalize_peak_db": float(normalize_peak_db),
        "normalize_scale": float(scale),
        "peak_before": float(peak),
    }


def apply_minimal_edit(raw_path: Path, output_path: Path, *, trim_enabled: bool, normalize_enabled: bool,
                       target_dbfs: float, silence_threshold: float = SILENCE_THRESHOLD,
                       silence_min_ms: int = SILENCE_MIN_MS, zero_cross_radius_ms: int = 10,
                       fade_ms: int = 10) -> dict[str, Any]:
    """GPU ``apply_minimal_edit`` (reference audio_edit.py:16-79): optional trim (no snap, no
    fades) + optional peak normalise + clip, written as PCM_16."""
    raw_path = Path(raw_path)
    output_path = Path(output_path)
    if raw_path.resolve() == output_path.resolve():
        raise ValueError("Output must be different from input.")
    torch = _torch()
    lib = _lib.load_library()
    q, sr = _wav.read_pcm16(raw_path)
    n = q.size
    target_peak = 10 ** (float(target_dbfs) / 20.0)
    prm = make_params(sr=sr, trim=1 if trim_enabled else 0, silence_threshold=float(silence_threshold),
                      min_silence_frames=int(sr * (int(silence_min_ms) / 1000.0)), snap_radius=-1,
                      normalize=1, clip=1, target_peak=float(target_peak), out_pcm16=1)
    # the peak is always measured (peak_before is reported even when normalisation is off);
    # the gain is only applied when requested
    trimmed = False
    normalized = False
    peak_before, gain = 0.0, 1.0
    out_q = np.zeros(0, np.int16)
    if n:
        qd = torch.from_numpy(q).cuda()
        x = torch.empty(n + 4, dtype=torch.float32, device="cuda")
        check(lib.vt_pcm16_decode(_ptr(qd), _ptr(x), n, _stream(torch)), "vt_pcm16_decode")
        if not normalize_enabled:
            # analyse with normalize=1 to obtain the peak, write with unit gain
            one = torch.zeros(1, dtype=torch.float32, device="cuda")  # peak_override 0 -> no gain applied
            r = post_process_device(x, [0, n], prm, peak_override=one)
        else:
            r = post_process_device(x, [0, n], prm)
        start, end = int(r.results[0, 0]), int(r.results[0, 1])
        trimmed = bool(trim_enabled and 0 <= start < end <= n)
        peak_before = float(r.results[0, 2])
        if normalize_enabled and peak_before > 0.0 and target_peak > 0.0:
            gain = float(r.results[0, 3])
            normalized = True
        out_q = r.out[:r.total].cpu().numpy()
    if normalized:
        peak_after = float(min(np.float32(peak_before) * np.float32(gain), np.float32(1.0)))
    else:
        peak_after = float(min(np.float32(peak_before), np.float32(1.0)))
    _wav.write_pcm16(output_path, out_q, sr)
    return {
        "trimmed": trimmed,
        "normalized": normalized,
        "target_dbfs": float(target_dbfs),
        "peak_before": peak_before,
        "peak_after": peak_after,
        "gain": gain,
    }


# ------------------------------------------------------------------------- resampling
RESAMPLE_ZEROS = 64        # zero crossings per side at the lower rate
RESAMPLE_ATT_DB = 120.0    # Kaiser stop-band attenuation
_resample_tables: dict = {}


def resample_out_length(n: int, orig_sr: int, target_sr: int) -> int:
    """librosa.resample's length rule (librosa/core/audio.py): ``int(np.ceil(n * (target_sr / orig_sr)))`` in float64 -
    including its quirk (22 050 samples at 22 050 Hz -> 24 001 at 24 kHz, because the product is 24000.000000000004)."""
    return int(np.ceil(int(n) * (float(target_sr) / orig_sr)))


def _resample_table(torch, orig_sr: int, target_sr: int, device):
    """Phase table of the Kaiser-windowed sinc (float64 design, fp32 table) - the numbers of oracle/resample_oracle.py
    restated here: the product path does not import the oracle."""
    import math
    g = math.gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g, int(orig_sr) // g
    key = (up, down, str(device))
    if key not in _resample_tables:
        r = max(up, down)
        half = RESAMPLE_ZEROS * r
        beta = 0.1102 * (RESAMPLE_ATT_DB - 8.7)
        dw = (RESAMPLE_ATT_DB - 8.0) / (2.285 * (2 * half))
        wc = math.pi / r - dw / 2.0
        n = np.arange(-half, half + 1, dtype=np.float64)
        h = (wc / math.pi) * np.sinc(wc / math.pi * n) * np.kaiser(2 * half + 1, beta)
        h /= h.sum()
        j0 = (half + up - 1) // up
        tab = np.zeros((up, 2 * j0 + 1), dtype=np.float64)
        idx = np.arange(up)[:, None] + (np.arange(-j0, j0 + 1) * up)[None, :]
        ok = np.abs(idx) <= half
        tab[ok] = up * h[(idx + half)[ok]]
        _resample_tables[key] = (up, down, torch.from_numpy(np.ascontiguousarray(tab.astype(np.float32))).to(device))
    return _resample_tables[key]


def resample_device(audio, seg_off, orig_sr: int, target_sr: int):
    """Resample every segment of a packed float32 CUDA buffer (segments ``seg_off``, host int64); returns the packed
    output tensor and its int64 offsets (host) - segment i has ``resample_out_length(len_i)`` samples."""
    torch = _torch()
    lib = _lib.load_library()
    if audio.dtype != torch.float32 or audio.dim() != 1 or not audio.is_cuda:
        raise ValueError("audio must be a 1-D float32 CUDA tensor")
    seg_np = np.ascontiguousarray(seg_off, dtype=np.int64)
    if seg_np.ndim != 1 or seg_np.size < 1 or np.any(np.diff(seg_np) < 0) or int(seg_np[-1]) > audio.numel():
        raise ValueError("seg_off must be non-decreasing and inside the audio buffer")
    lens = np.diff(seg_np)
    out_off = np.concatenate([[0], np.cumsum([resample_out_length(int(n), orig_sr, target_sr) for n in lens])]).astype(np.int64)
    dev = audio.device
    with torch.cuda.device(dev):
        up, down, tab = _resample_table(torch, orig_sr, target_sr, dev)
        out = torch.empty(max(int(out_off[-1]), 4), dtype=torch.float32, device=dev)
        if len(lens):
            src = audio.contiguous()
            off_in_dev, off_out_dev = torch.from_numpy(seg_np).to(dev), torch.from_numpy(out_off).to(dev)   # named: they must outlive the call
            check(lib.vt_resample(_ptr(src), _ptr(off_in_dev), _ptr(off_out_dev), len(lens), int(np.diff(out_off).max()),
                                  up, down, _ptr(tab), int(tab.shape[1]), _ptr(out), _stream(torch)), "vt_resample")
    return out, out_off


def _resample_audio(audio: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """GPU ``_resample_audio`` (reference tts_pipeline.py:100-111): same signature and shapes - 1-D mono, or
    ``[n, channels]`` resampled per channel and cut to the shortest - with the output length of ``librosa.resample``.
    The impulse response is this package's 120 dB Kaiser sinc, not soxr's (absent here): values are not bit-pinned."""
    if orig_sr == target_sr:
        return audio
    torch = _torch()
    a = np.asarray(audio, dtype=np.float32)
    if a.ndim == 1:
        if a.size == 0:
            return np.zeros(0, dtype=np.float32)
        out, off = resample_device(torch.from_numpy(np.ascontiguousarray(a)).cuda(), [0, a.size], orig_sr, target_sr)
        return out[: int(off[-1])].cpu().numpy()
    n, ch = a.shape[0], a.shape[1]
    if ch == 0 or n == 0:
        return np.zeros(0, dtype=np.float32)                                   # tts_pipeline.py:108-110
    flat = np.ascontiguousarray(a.T).reshape(-1)                               # channel-major: one segment per channel
    out, off = resample_device(torch.from_numpy(flat).cuda(), np.arange(ch + 1, dtype=np.int64) * n, orig_sr, target_sr)
    y = out[: int(off[-1])].cpu().numpy().reshape(ch, -1)
    return np.ascontiguousarray(y.T)


def pcm16_encode(audio: np.ndarray) -> np.ndarray:
    """float32 -> PCM_16 codes on the GPU (``lrintf(x * 32767)``, the libsndfile default the
    reference writes with: tts_backends/chatterbox_runner.py:152, tts_pipeline.py:409)."""
    torch = _torch()
    lib = _lib.load_library()
    x = _as_f32(audio)
    if x.size == 0:
        return np.zeros(0, np.int16)
    xd = torch.from_numpy(x).cuda()
    q = torch.empty(x.size, dtype=torch.int16, device="cuda")
    check(lib.vt_pcm16_encode(_ptr(xd), _ptr(q), x.size, _stream(torch)), "vt_pcm16_encode")
    return q.cpu().numpy()


def rms(audio: np.ndarray) -> float:
    """``sqrt(mean(x.astype(float64) ** 2))`` on the GPU - the helper the reference uses to validate
    clips (tts_backends/cosyvoice_backend.py:103, tests/test_qwen3_runner.py:58); 0.0 for an empty array."""
    return float(rms_segments(audio, [0, int(np.asarray(audio).size)])[0])


def rms_segments(audio: np.ndarray, seg_off: Sequence[int]) -> np.ndarray:
    """Per-segment RMS (float64) of a flat float32 buffer; segment ``i`` is ``audio[seg_off[i]:seg_off[i+1]]``."""
    torch = _torch()
    lib = _lib.load_library()
    x = _as_f32(audio)
    off = np.asarray(seg_off, dtype=np.int64)
    n_seg = off.size - 1
    if n_seg <= 0:
        return np.zeros(0, np.float64)
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda() if x.size else torch.zeros(4, dtype=torch.float32, device="cuda")
    od = torch.from_numpy(off).cuda()
    out = torch.empty(n_seg, dtype=torch.float64, device="cuda")
    ws = torch.empty(n_seg * 64, dtype=torch.float64, device="cuda")
    check(lib.vt_rms(_ptr(xd), _ptr(od), n_seg, _ptr(out), _ptr(ws), ws.numel() * 8, _stream(torch)), "vt_rms")
    return out.cpu().numpy()
