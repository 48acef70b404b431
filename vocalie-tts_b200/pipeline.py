"""Mel -> finished audio on the GPU: the B200 replacement of the reference's serial chunk loop
tail (backend/shared/tts_pipeline.py:353-409 after synthesis, plus the post-processing of
tts_pipeline.py:162-274 / audio_edit.py:16-79), batched over the independent chunks of a job.

``VocoderPipeline.run`` is the call a user of this package makes: host mel buffers in, host audio
out.  ``run_device`` is the same with device-resident tensors (no copies) for callers that
already hold mels on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import post as _post
from .hift import HiFTVocoder, SAMPLES_PER_FRAME, S3GEN_SR, N_MEL, _torch


@dataclass
class JobResult:
    audio: "object"              # float32 (or int16) tensor/array holding the stitched job audio
    total_samples: int
    segments: Optional[np.ndarray]   # [n_chunks, 8] per-chunk start,end,peak,scale,dst,len,peak_used,-
    sr: int = S3GEN_SR


class VocoderPipeline:
    """HiFT vocoder + per-chunk post-processing + gap stitching on one GPU."""

    def __init__(self, vocoder: HiFTVocoder, *, chunk_gap_ms: int = 250, trim_silence: bool = True,
                 normalize: bool = True, target_dbfs: float = -1.0, fade_ms: int = 10,
                 zero_cross_radius_ms: int = 10, silence_threshold: float = _post.SILENCE_THRESHOLD,
                 silence_min_ms: int = _post.SILENCE_MIN_MS, out_pcm16: bool = False):
        self.voc = vocoder
        self.sr = S3GEN_SR
        self.opts = dict(chunk_gap_ms=int(chunk_gap_ms), trim_silence=bool(trim_silence), normalize=bool(normalize),
                         target_dbfs=float(target_dbfs), fade_ms=int(fade_ms), zero_cross_radius_ms=int(zero_cross_radius_ms),
                         silence_threshold=float(silence_threshold), silence_min_ms=int(silence_min_ms),
                         out_pcm16=bool(out_pcm16))
        self.stitch_head, self.stitch_tail = 1, 1     # single-GPU job; a shard of a multi-GPU job sets these
        self._host_in = None
        self._host_out = None
        self._dev_in = None
        self._wav = None
        self._out = None
        self.last_launches = 0
        self._slots = None            # submit()/collect(): double-buffered device output + pinned host staging
        self._copy_stream = None
        self._h2d_done = None
        self._n_submitted = 0

    def set_shard(self, chunk_ids, n_total_chunks: int) -> None:
        """Declare that this pipeline holds ``chunk_ids`` (sorted job-order indices) of a job sharded
        over several GPUs: only the job's first chunk skips the fade-in, only its last chunk skips the
        fade-out and the trailing gap (see distributed.py)."""
        from .distributed import stitch_flags
        self.stitch_head, self.stitch_tail = stitch_flags(chunk_ids, n_total_chunks)

    def post_params(self, n_chunks: int):
        o = self.opts
        sr = self.sr
        gap_on = o["chunk_gap_ms"] > 0 and (n_chunks > 1 or not (self.stitch_head and self.stitch_tail))
        fade = _post._ms_to_frames(sr, o["fade_ms"])
        return _post.make_params(
            sr=sr, trim=1 if o["trim_silence"] else 0, silence_threshold=o["silence_threshold"],
            min_silence_frames=_post._ms_to_frames(sr, o["silence_min_ms"]),
            snap_radius=_post._ms_to_frames(sr, o["zero_cross_radius_ms"]) if o["trim_silence"] else -1,
            fade_in_frames=fade if gap_on else 0, fade_out_frames=fade if gap_on else 0, stitch=1,
            gap_frames=_post._ms_to_frames(sr, o["chunk_gap_ms"]) if gap_on else 0,
            normalize=1 if o["normalize"] else 0, target_peak=float(10 ** (o["target_dbfs"] / 20.0)), concat=1,
            out_pcm16=1 if o["out_pcm16"] else 0, stitch_head=self.stitch_head, stitch_tail=self.stitch_tail)

    def run_device(self, mel, T, *, f0=None, phase_vec=None, noise=None, seed: int = 0, read_back: bool = False) -> JobResult:
        """``mel``: float32 CUDA [sum_T, 80]; ``T``: int32 frames per chunk (host)."""
        torch = _torch()
        T = np.ascontiguousarray(T, dtype=np.int32)
        n = int(T.astype(np.int64).sum()) * SAMPLES_PER_FRAME
        if self._wav is None or self._wav.numel() < n + 4:
            self._wav = torch.empty(n + 4, dtype=torch.float32, device=self.voc.device)
        wav = self.voc.forward_packed(mel, T, f0=f0, phase_vec=phase_vec, noise=noise, seed=seed, out=self._wav)
        seg_off = np.concatenate([[0], np.cumsum(T.astype(np.int64) * SAMPLES_PER_FRAME)])
        prm = self.post_params(len(T))
        cap = n + len(T) * int(prm.gap_frames)
        odt = torch.int16 if prm.out_pcm16 else torch.float32
        if self._out is None or self._out.numel() < max(cap, 4) or self._out.dtype != odt:
            self._out = torch.empty(max(cap, 4), dtype=odt, device=self.voc.device)
        r = _post.post_process_device(wav, seg_off, prm, out=self._out, read_back=read_back)
        # the library's per-thread launch counter is reset by vt_hift_forward and keeps counting through the post calls
        self.last_launches = _post.last_launch_count()
        self._last_post = r
        return JobResult(r.out, r.total if r.total is not None else cap, r.results)

    def run(self, mel_host, T, *, seed: int = 0) -> JobResult:
        """Host in / host out: ``mel_host`` float32 [sum_T, 80] (numpy or CPU tensor; pinned staging is
        managed here), returns the stitched job audio as a numpy array of exactly the right length."""
        torch = _torch()
        T = np.ascontiguousarray(T, dtype=np.int32)
        total_T = int(T.astype(np.int64).sum())
        src = torch.as_tensor(mel_host, dtype=torch.float32).reshape(total_T, N_MEL)
        if self._host_in is None or self._host_in.numel() < src.numel():
            self._host_in = torch.empty(src.numel(), dtype=torch.float32).pin_memory()
            self._dev_in = torch.empty(src.numel(), dtype=torch.float32, device=self.voc.device)
        hin = self._host_in[: src.numel()].view(total_T, N_MEL)
        if src.data_ptr() != hin.data_ptr():
            hin.copy_(src)
        dev = self._dev_in[: src.numel()].view(total_T, N_MEL)
        dev.copy_(hin, non_blocking=True)
        res = self.run_device(dev, T, seed=seed, read_back=True)
        n = int(res.total_samples)
        if self._host_out is None or self._host_out.numel() < n or self._host_out.dtype != res.audio.dtype:
            self._host_out = torch.empty(max(n, 4), dtype=res.audio.dtype).pin_memory()
        self._host_out[:n].copy_(res.audio[:n], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return JobResult(self._host_out[:n].numpy(), n, res.segments)

    # ---- pipelined serving loop: submit(job k+1) is enqueued while job k's audio is still crossing PCIe
    def submit(self, mel_host, T, *, seed: int = 0) -> int:
        """Enqueue one job (pinned host mels -> device -> HiFT -> post) without waiting for it and start the
        device-to-host copy of its audio on a second stream; returns a ticket for ``collect``.  At most two jobs may
        be in flight (two output slots).  The input is staged exactly like ``run``; a caller that fills
        ``pinned_input()`` in place must not overwrite it before the NEXT submit/collect returns."""
        torch = _torch()
        T = np.ascontiguousarray(T, dtype=np.int32)
        total_T = int(T.astype(np.int64).sum())
        n_cap = total_T * SAMPLES_PER_FRAME + len(T) * int(self.post_params(len(T)).gap_frames)
        odt = torch.int16 if self.opts["out_pcm16"] else torch.float32
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.voc.device)
        if self._slots is None:
            self._slots = [dict(dev=None, host=None, tot=torch.zeros(1, dtype=torch.int64).pin_memory(), res=None, n_seg=0,
                                done=torch.cuda.Event(), busy=False) for _ in range(2)]
        ticket = self._n_submitted
        slot = self._slots[ticket & 1]
        if slot["busy"]:
            raise RuntimeError("two jobs are already in flight: collect() one first")
        if slot["dev"] is None or slot["dev"].numel() < max(n_cap, 4) or slot["dev"].dtype != odt:
            # a free slot grows on its own: the other one may still be in flight
            slot["dev"] = torch.empty(max(n_cap, 4), dtype=odt, device=self.voc.device)
            slot["host"] = torch.empty(max(n_cap, 4), dtype=odt).pin_memory()
        src = torch.as_tensor(mel_host, dtype=torch.float32).reshape(total_T, N_MEL)
        if self._h2d_done is not None:
            self._h2d_done.synchronize()      # the previous job's input has left the pinned staging buffer
        hin = self.pinned_input(total_T)
        if src.data_ptr() != hin.data_ptr():
            hin.copy_(src)
        dev = self._dev_in[: src.numel()].view(total_T, N_MEL)
        cur = torch.cuda.current_stream()
        dev.copy_(hin, non_blocking=True)
        self._h2d_done = torch.cuda.Event()
        self._h2d_done.record(cur)
        saved_out = self._out
        self._out = slot["dev"]
        try:
            self.run_device(dev, T, seed=seed, read_back=False)
        finally:
            self._out = saved_out
        r = self._last_post
        ready = torch.cuda.Event()
        ready.record(cur)
        if slot["res"] is None or slot["res"].shape != r.results_dev.shape:
            slot["res"] = torch.empty(r.results_dev.shape, dtype=r.results_dev.dtype).pin_memory()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            slot["tot"].copy_(r.total_dev, non_blocking=True)
            slot["res"].copy_(r.results_dev, non_blocking=True)
            slot["host"][:n_cap].copy_(slot["dev"][:n_cap], non_blocking=True)   # worst-case length: the total is not known yet
            slot["done"].record(self._copy_stream)
        # the compute stream must not reuse the slot before its copy has left (it is reused two submits later)
        slot.update(n_seg=len(T), busy=True, bytes=n_cap * slot["dev"].element_size(), keep=r)   # keep: the copy reads r's tensors
        self._n_submitted += 1
        return ticket

    def collect(self, ticket: int) -> JobResult:
        """Wait for a submitted job; returns its stitched audio as a numpy view of the pinned slot (valid until the
        slot is reused by the second submit after this one)."""
        torch = _torch()
        slot = self._slots[ticket & 1]
        if not slot["busy"]:
            raise RuntimeError("ticket is not in flight")
        slot["done"].synchronize()
        torch.cuda.current_stream().wait_event(slot["done"])
        slot["busy"] = False
        n = int(slot["tot"].item())
        self.last_d2h_bytes = slot["bytes"]
        return JobResult(slot["host"][:n].numpy(), n, slot["res"].numpy()[: slot["n_seg"]].copy())

    def pinned_input(self, total_T: int):
        """A pinned host buffer [total_T, 80] callers can fill in place to skip the staging copy."""
        torch = _torch()
        n = total_T * N_MEL
        if self._host_in is None or self._host_in.numel() < n:
            self._host_in = torch.empty(n, dtype=torch.float32).pin_memory()
            self._dev_in = torch.empty(n, dtype=torch.float32, device=self.voc.device)
        return self._host_in[:n].view(total_T, N_MEL)
