"""Drop-in backend at the reference's plugin boundary.

``ChatterboxB200Backend`` honours the contract of ``TTSBackend`` (reference
tts_backends/base.py:50-217) as specialised by ``ChatterboxBackend``
(tts_backends/chatterbox_backend.py:16-192) and consumed by ``run_tts_pipeline``
(backend/shared/tts_pipeline.py:366-371): same class attributes, same ``synthesize`` /
``synthesize_chunk`` signatures and return shapes, the single error type
``BackendUnavailableError``.  What changes is everything below the boundary: instead of one
subprocess + model load per chunk running HiFT on CPU, the mel of every chunk goes through the
resident B200 vocoder (``hift.HiFTVocoder``) and, for whole jobs, through the fused
post-processing of ``pipeline.VocoderPipeline``.

Out of scope: text -> speech tokens (T3) and tokens -> mel (S3Gen flow) stay
upstream's; they are injected as a ``mel_provider`` callable
``(text, voice_ref_path=None, lang=None, **params) -> mel [80, T]`` (or a dict with ``mel`` and
optionally ``f0``).  ``get_backend`` instantiates a new object per call
(tts_backends/__init__.py:51-65), so the engine state lives at class level.
"""
from __future__ import annotations

import itertools
import os
import threading
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from . import wav as _wav
from .errors import BackendUnavailableError

try:  # inside a Vocalie-TTS checkout: extend the reference's own ABC so the registry is shared
    from tts_backends.base import TTSBackend, ParamSpec  # type: ignore
    _IN_REFERENCE = True
except Exception:  # standalone: a minimal mirror of the same interface
    _IN_REFERENCE = False
    from abc import ABC, abstractmethod
    from dataclasses import dataclass

    @dataclass(frozen=True)
    class ParamSpec:  # tts_backends/base.py:35-47
        key: str
        type: str
        default: Any
        min: Optional[float] = None
        max: Optional[float] = None
        step: Optional[float] = None
        choices: Optional[List[Any]] = None
        label: Optional[str] = None
        help: Optional[str] = None
        visible_if: Optional[Dict[str, Any]] = None
        serialize_scope: str = "engine"

    class TTSBackend(ABC):  # tts_backends/base.py:50-217 (the members the hot path touches)
        _REGISTRY: Dict[str, type] = {}
        id: str
        display_name: str
        supports_ref_audio: bool = False
        uses_internal_voices: bool = False
        supports_inter_chunk_gap: bool = False

        def __init_subclass__(cls, **kwargs) -> None:
            super().__init_subclass__(**kwargs)
            if getattr(cls, "id", None) and not getattr(cls, "__abstractmethods__", None):
                TTSBackend._REGISTRY[cls.id] = cls

        @classmethod
        def is_available(cls) -> bool:
            return True

        @classmethod
        def unavailable_reason(cls) -> Optional[str]:
            return None

        def supported_languages(self) -> List[str]:
            return []

        def map_language(self, bcp47: Optional[str]) -> Optional[str]:
            return bcp47 if (self.supported_languages() and bcp47) else None

        @abstractmethod
        def synthesize(self, script: str, out_path: str, voice_ref_path: Optional[str] = None,
                       lang: Optional[str] = None, **params: Any) -> Dict[str, Any]:
            raise NotImplementedError


SR = 24000   # backend/shared/tts_pipeline.py:26 TARGET_SR == upstream S3GEN_SR
_LANGS = ["fr-FR", "en-US", "en-GB", "es-ES", "de-DE", "it-IT", "pt-PT", "nl-NL"]


class ChatterboxB200Backend(TTSBackend):
    """Chatterbox with the mel -> waveform tail on a B200 (re-uses id ``chatterbox`` so that
    importing this module replaces the stock backend in ``TTSBackend._REGISTRY``)."""

    id = "chatterbox"
    display_name = "Chatterbox (B200 vocoder)"
    supports_ref_audio = True
    uses_internal_voices = False
    supports_inter_chunk_gap = True

    # class-level engine: one resident vocoder per process, guarded for the reference's job threads
    # (up to MAX_CONCURRENT_JOBS = 2, backend/config.py:11)
    _lock = threading.Lock()
    _vocoder = None
    _mel_provider: Optional[Callable[..., Any]] = None
    _why_unavailable: Optional[str] = "not configured: call ChatterboxB200Backend.configure(...)"

    # Upstream draws fresh SineGen randomness (phase_vec, noise) from torch's global RNG on every call; the in-kernel
    # Philox stream is keyed on (seed, sequence, sample), so a caller that passes no seed gets a new one per call
    # (process-random base + counter) instead of the same phases and noise for every chunk of every job.
    _seed_base = int.from_bytes(os.urandom(7), "little")
    _seed_counter = itertools.count()

    @classmethod
    def _next_seed(cls) -> int:
        return (cls._seed_base + 0x9E3779B97F4A7C15 * next(cls._seed_counter)) & (2 ** 63 - 1)

    # ------------------------------------------------------------------ configuration
    @classmethod
    def configure(cls, *, state_dict=None, vocoder=None, mel_provider: Callable[..., Any], operand: str = "fp16") -> None:
        """Install the resident vocoder (from an upstream HiFT ``state_dict`` or an existing
        ``HiFTVocoder``) and the upstream text->mel callable."""
        with cls._lock:
            try:
                if vocoder is None:
                    if state_dict is None:
                        raise ValueError("configure() needs a HiFT state_dict or a HiFTVocoder")
                    from .hift import HiFTVocoder
                    vocoder = HiFTVocoder(state_dict, operand=operand)
            except BackendUnavailableError as exc:
                cls._vocoder, cls._mel_provider, cls._why_unavailable = None, None, str(exc)
                raise
            cls._vocoder = vocoder
            cls._mel_provider = mel_provider
            cls._why_unavailable = None

    @classmethod
    def reset(cls) -> None:
        with cls._lock:
            cls._vocoder = None
            cls._mel_provider = None
            cls._why_unavailable = "not configured: call ChatterboxB200Backend.configure(...)"

    @classmethod
    def is_available(cls) -> bool:
        return cls._vocoder is not None and cls._mel_provider is not None

    @classmethod
    def unavailable_reason(cls) -> Optional[str]:
        return None if cls.is_available() else cls._why_unavailable

    @classmethod
    def engine_variants(cls) -> List[Dict[str, str]]:
        return [{"id": "chatterbox_native", "label": "Chatterbox (native multilang)"},
                {"id": "chatterbox_finetune_fr", "label": "Chatterbox (FR fine-tune)"}]

    def supported_languages(self) -> List[str]:
        return list(_LANGS)

    def default_language(self) -> str:
        return "fr-FR"

    def map_language(self, bcp47: Optional[str]) -> Optional[str]:
        if not bcp47:
            return "fr"
        return bcp47.split("-")[0]

    def params_schema(self) -> Dict[str, ParamSpec]:
        # same knobs as tts_backends/chatterbox_backend.py:53-113; they are forwarded to the mel provider
        return {
            "chatterbox_mode": ParamSpec("chatterbox_mode", "choice", "fr_finetune",
                                         choices=[("FR fine-tuné (spécialisé)", "fr_finetune"), ("Multilangue", "multilang")]),
            "multilang_cfg_weight": ParamSpec("multilang_cfg_weight", "float", 0.5, 0.0, 1.0, 0.05),
            "exaggeration": ParamSpec("exaggeration", "float", 0.5, 0.0, 2.0, 0.05),
            "cfg_weight": ParamSpec("cfg_weight", "float", 0.6, 0.0, 1.0, 0.05),
            "temperature": ParamSpec("temperature", "float", 0.5, 0.05, 2.0, 0.05),
            "repetition_penalty": ParamSpec("repetition_penalty", "float", 1.35, 1.0, 2.0, 0.05),
        }

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _engine_params(params: Dict[str, Any]) -> Dict[str, Any]:
        """The subset the reference forwards to its runner (chatterbox_backend.py:195-214); unknown keys
        (voice, model_id, inter_chunk_gap_ms ... injected by run_tts_job) are ignored."""
        return {
            "tts_model_mode": params.get("tts_model_mode", params.get("chatterbox_mode", "fr_finetune")),
            "multilang_cfg_weight": params.get("multilang_cfg_weight", 0.5),
            "exaggeration": params.get("exaggeration", 0.5),
            "cfg_weight": params.get("cfg_weight", 0.6),
            "temperature": params.get("temperature", 0.5),
            "repetition_penalty": params.get("repetition_penalty", 1.35),
        }

    @classmethod
    def _engine(cls):
        if not cls.is_available():
            raise BackendUnavailableError(cls._why_unavailable or "chatterbox B200 backend unavailable")
        return cls._vocoder, cls._mel_provider

    @staticmethod
    def _split_provider_result(res):
        if isinstance(res, dict):
            return res["mel"], res.get("f0")
        return res, None

    def _mels_for(self, texts: Sequence[str], voice_ref_path, lang, params):
        _, provider = self._engine()
        ep = self._engine_params(params)
        mels, f0s = [], []
        for text in texts:
            if not str(text or "").strip():
                raise ValueError("Texte vide.")   # same complaint as the pipeline for empty chunks
            try:
                mel, f0 = self._split_provider_result(provider(text, voice_ref_path=voice_ref_path, lang=lang, **ep))
            except (BackendUnavailableError, ValueError):
                raise
            except Exception as exc:  # everything below the boundary surfaces as one error type
                raise BackendUnavailableError(f"mel provider failed: {exc}") from exc
            mels.append(mel)
            f0s.append(f0)
        return mels, f0s

    # ------------------------------------------------------------------ the boundary
    def synthesize_chunks(self, texts: Sequence[str], *, voice_ref_path: Optional[str] = None,
                          lang: Optional[str] = None, **params: Any):
        """Batched ``synthesize_chunk``: one vocoder launch sequence for all chunks of a job.
        Returns a list of ``(audio float32 mono, sr, meta)``."""
        import torch
        voc, _ = self._engine()
        mels, f0s = self._mels_for(texts, voice_ref_path, lang, params)
        mels = [torch.as_tensor(np.asarray(m) if not hasattr(m, "shape") else m) for m in mels]
        f0 = None
        if all(f is not None for f in f0s) and f0s:
            f0 = [torch.as_tensor(f) for f in f0s]
        try:
            with self._lock:
                seed = int(params["seed"]) if params.get("seed") is not None else self._next_seed()
                wavs = voc.inference(mels, f0=f0, seed=seed)
                out = [w.cpu().numpy() for w in wavs]
        except BackendUnavailableError:
            raise
        except RuntimeError as exc:
            raise BackendUnavailableError(f"B200 vocoder failed: {exc}") from exc
        return [(np.asarray(a, dtype=np.float32), SR, {"retry": False, "backend_id": self.id, "backend_lang": lang})
                for a in out]

    def synthesize_chunk(self, text: str, *, voice_ref_path: Optional[str] = None, lang: Optional[str] = None,
                         **params: Any):
        """``(np.ndarray float32 mono, sr, meta)`` - tts_backends/base.py:190-217."""
        return self.synthesize_chunks([text], voice_ref_path=voice_ref_path, lang=lang, **params)[0]

    def synthesize(self, script: str, out_path: str, voice_ref_path: Optional[str] = None,
                   lang: Optional[str] = None, **params: Any) -> Dict[str, Any]:
        """Writes a PCM_16 WAV at ``out_path`` (the runner's wire format,
        tts_backends/chatterbox_runner.py:152) and returns the reference's meta dict
        (chatterbox_backend.py:163-174).  The file is finished ON THE DEVICE: the vocoder's float32 waveform goes
        through the fused PCM_16 writer behind a device-written RIFF header and leaves the GPU in one copy
        (``post.WavImage``) - no float32 round trip over PCIe, no host-side encode."""
        from . import post as _post
        torch = _post._torch()          # no CUDA device -> BackendUnavailableError (no CPU fallback below the boundary)
        voc, _ = self._engine()
        mels, f0s = self._mels_for([script], voice_ref_path, lang, params)
        mel = torch.as_tensor(np.asarray(mels[0]) if not hasattr(mels[0], "shape") else mels[0])
        f0 = None if f0s[0] is None else [torch.as_tensor(f0s[0])]
        try:
            with self._lock:
                seed = int(params["seed"]) if params.get("seed") is not None else self._next_seed()
                with torch.cuda.device(voc.device):
                    wav = voc.inference([mel], f0=f0, seed=seed)[0]
                    n = int(wav.numel())
                    img = _post.WavImage(n, voc.device)
                    prm = _post.make_params(sr=SR, concat=1, out_pcm16=1)
                    _post.post_process_device(wav, [0, n], prm, out=img.samples, read_back=False)
                    img.finish(SR, total=n)
                    img.to_file(out_path, n)
        except BackendUnavailableError:
            raise
        except RuntimeError as exc:
            raise BackendUnavailableError(f"B200 vocoder failed: {exc}") from exc
        return {"backend_id": self.id, "backend_lang": lang, "out_path": str(out_path),
                "duration_s": float(n) / float(SR), "retry": False}


def shard_chunks(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of chunks to ranks: cost is proportional to
    the mel length; returns, per rank, the chunk indices in their original (output) order."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(p) for p in parts]
