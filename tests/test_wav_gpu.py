"""GPU tests of the fused WAV path: PCM_16 encode + RIFF header on the device, one D2H copy.  Files
must be BYTE-identical to the host writer (wav.write_pcm16 = Python's wave module, the container libsndfile also writes
for sf.write(path, x, sr)) fed with the lrintf(x * 32767) codes of the same waveform."""
import io
import json
import wave
from pathlib import Path

import numpy as np
import pytest

from oracle import hift_oracle as H
from oracle import post_oracle as po

pytestmark = pytest.mark.gpu
SR = 24000


@pytest.fixture(scope="module")
def voc():
    import torch
    assert torch.cuda.is_available()
    from vocalie_tts_b200.hift import HiFTVocoder
    return HiFTVocoder(H.make_state_dict(0, "unit"), operand="fp16")


def _provider(text, voice_ref_path=None, lang=None, **params):
    T = 10 + len(text)
    return {"mel": H.synth_mel(T, 13, len(text)), "f0": H.synth_f0(T, 13, len(text))}


def test_synthesize_writes_the_same_bytes_as_the_host_writer(voc, tmp_path):
    from vocalie_tts_b200 import backend as B, wav
    B.ChatterboxB200Backend.configure(vocoder=voc, mel_provider=_provider)
    try:
        be = B.ChatterboxB200Backend()
        for i, text in enumerate(["Salut.", "Une phrase un peu plus longue pour un second fichier."]):
            out = tmp_path / f"dev{i}.wav"
            meta = be.synthesize(text, str(out), lang="fr-FR", seed=5 + i)
            audio, sr, _ = be.synthesize_chunk(text, lang="fr-FR", seed=5 + i)
            ref = tmp_path / f"host{i}.wav"
            wav.write_pcm16(ref, po.pcm16_encode(audio), sr)
            assert out.read_bytes() == ref.read_bytes(), text
            assert meta["duration_s"] == audio.size / SR and Path(meta["out_path"]) == out
            with wave.open(str(out), "rb") as w:
                assert (w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()) == (SR, 1, 2, audio.size)
    finally:
        B.ChatterboxB200Backend.reset()


def test_job_wav_with_device_side_length(voc, tmp_path):
    """A trimmed job: the sample count in the header comes from device memory (the writer's total)."""
    import torch
    from vocalie_tts_b200 import wav
    from vocalie_tts_b200.pipeline import VocoderPipeline
    Ts = np.array([20, 35, 9], np.int32)
    mel = torch.cat([H.synth_mel(int(T), 12, b).t() for b, T in enumerate(Ts)]).contiguous().numpy()
    for gran in ("job", "chunk"):
        pipe = VocoderPipeline(voc, chunk_gap_ms=250, out_pcm16=True, granularity=gran)
        r = pipe.run(mel, Ts, seed=3)
        got = tmp_path / f"{gran}.wav"
        rw = pipe.run_wav(mel, Ts, got, seed=3)
        want = tmp_path / f"{gran}_host.wav"
        wav.write_pcm16(want, r.audio, SR)
        assert rw.total_samples == r.total_samples
        assert got.read_bytes() == want.read_bytes(), gran


def test_resident_worker_serves_real_wavs(voc, tmp_path):
    from vocalie_tts_b200 import backend as B, wav
    from vocalie_tts_b200.worker import ResidentWorker
    B.ChatterboxB200Backend.configure(vocoder=voc, mel_provider=_provider)
    try:
        w = ResidentWorker()
        reqs = [{"text": f"Morceau numero {i}.", "out_path": str(tmp_path / f"w{i}.wav"), "lang": "fr", "seed": 40 + i} for i in range(3)]
        fout = io.StringIO()
        w.serve_stream(io.StringIO("".join(json.dumps(r) + "\n" for r in reqs)), fout)
        resps = [json.loads(l) for l in fout.getvalue().splitlines()]
        assert [r["ok"] for r in resps] == [True] * 3 and w.served == 3
        be = B.ChatterboxB200Backend()
        for i, (req, resp) in enumerate(zip(reqs, resps)):
            audio, sr, _ = be.synthesize_chunk(req["text"], lang="fr", seed=40 + i)
            ref = tmp_path / f"ref{i}.wav"
            wav.write_pcm16(ref, po.pcm16_encode(audio), sr)
            assert Path(resp["out_path"]).read_bytes() == ref.read_bytes()
            assert abs(resp["duration_s"] - audio.size / SR) < 1e-12
    finally:
        B.ChatterboxB200Backend.reset()
