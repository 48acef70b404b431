"""The drop-in boundary exercised against the REAL reference, not the standalone mirror: in a fresh interpreter
with /root/reference on sys.path, importing ``vocalie_tts_b200.backend`` must (a) subclass the reference's own
``TTSBackend`` / raise its own ``BackendUnavailableError``, (b) replace ``_REGISTRY["chatterbox"]``, and (c) serve
the reference's unmodified ``run_tts_pipeline`` (backend/shared/tts_pipeline.py:292-430) end to end - chunking,
``synthesize_chunk`` per chunk, ``_apply_inter_chunk_gap``, ``sf.write`` - with a stub vocoder standing in for the
GPU (CPU test).  Skipped where /root/reference does not exist (the GPU box)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")

SCRIPT = r'''
import json, sys, types, wave
from pathlib import Path
import numpy as np
ROOT, REF, OUT = Path(sys.argv[1]), Path(sys.argv[2]), Path(sys.argv[3])
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(REF))
from oracle import sf_stub
sys.modules.setdefault("librosa", types.ModuleType("librosa"))
sys.modules["soundfile"] = sf_stub.make_module()
import tts_backends                                            # the reference's registry, stock backends registered
from tts_backends.base import TTSBackend, BackendUnavailableError
stock = TTSBackend._REGISTRY["chatterbox"]
import vocalie_tts_b200.backend as B                           # re-registers id "chatterbox"
import vocalie_tts_b200.errors as E
assert B._IN_REFERENCE and issubclass(B.ChatterboxB200Backend, TTSBackend)
assert E.BackendUnavailableError is BackendUnavailableError
assert TTSBackend._REGISTRY["chatterbox"] is B.ChatterboxB200Backend and stock is not B.ChatterboxB200Backend
be = tts_backends.get_backend("chatterbox")
assert isinstance(be, B.ChatterboxB200Backend) and be.supports_inter_chunk_gap
import backend.shared.tts_pipeline as tp

# unconfigured -> the reference's own availability gate fires with its own error type
try:
    tp.run_tts_pipeline({"tts_backend": "chatterbox", "script": "Bonjour.", "out_path": str(OUT / "x.wav")})
    raise SystemExit("expected BackendUnavailableError")
except BackendUnavailableError:
    pass

import torch
class StubVocoder:                                             # stands in for hift.HiFTVocoder on a CPU-only box
    calls = 0
    def inference(self, mels, f0=None, seed=0):
        StubVocoder.calls += 1
        return [0.25 * torch.sin(torch.arange(480 * m.shape[-1], dtype=torch.float32) * 0.05) for m in mels]
texts = []
def provider(text, voice_ref_path=None, lang=None, **params):
    texts.append((text, lang, dict(params)))
    return torch.zeros(80, 20 + len(text) % 7)
B.ChatterboxB200Backend.configure(vocoder=StubVocoder(), mel_provider=provider)
script = ("Bonjour tout le monde, ceci est un premier paragraphe assez long pour former un bloc. "
          "Voici une deuxieme phrase qui continue le propos avec suffisamment de mots pour etre decoupee. "
          "Et enfin une troisieme phrase, la derniere, qui termine ce petit texte de demonstration.")
res = tp.run_tts_pipeline({"tts_backend": "chatterbox", "script": script, "out_path": str(OUT / "job.wav"), "lang": "fr-FR",
                           "engine_params": {"temperature": 0.7, "voice": "ignored"}, "inter_chunk_gap_ms": 250,
                           "chunk_settings": {"min_words_per_chunk": 4, "max_words_without_terminator": 12, "max_est_seconds_per_chunk": 4.0}})
m = res.meta
assert m["backend_id"] == "chatterbox" and m["chunks"] == len(texts) >= 2 and StubVocoder.calls == len(texts)
assert m["inter_chunk_gap_applied"] is True and m["sr"] == 24000
assert all(t[2]["temperature"] == 0.7 and "voice" not in t[2] for t in texts)
with wave.open(res.out_path, "rb") as w:
    n = w.getnframes(); assert (w.getframerate(), w.getnchannels(), w.getsampwidth()) == (24000, 1, 2)
want = sum(480 * (20 + len(t[0]) % 7) for t in texts) + (len(texts) - 1) * 6000
assert n == want, (n, want)
# synthesize() encodes the WAV on the GPU: on this CPU-only box it must fail LOUDLY with the reference's own error
# type - there is no CPU fallback below the boundary
if not torch.cuda.is_available():
    try:
        be.synthesize("Salut.", str(OUT / "one.wav"), lang="fr-FR")
        raise SystemExit("expected BackendUnavailableError without a GPU")
    except BackendUnavailableError:
        pass
print(json.dumps({"ok": True, "chunks": m["chunks"], "frames": n}))
'''


@pytest.mark.skipif(not REF.exists(), reason="/root/reference is not present (GPU box)")
def test_b200_backend_serves_the_real_reference_pipeline(tmp_path):
    r = subprocess.run([sys.executable, "-c", SCRIPT, str(ROOT), str(REF), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ok"] and out["chunks"] >= 2
