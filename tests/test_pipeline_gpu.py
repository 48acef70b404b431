"""GPU tests of the job pipeline (mel -> HiFT -> per-chunk post -> stitched audio) and of the
drop-in backend, against the oracles."""
import wave

import numpy as np
import pytest

from oracle import hift_oracle as H
from oracle import post_oracle as po

pytestmark = pytest.mark.gpu
SR = 24000


@pytest.fixture(scope="module")
def setup():
    import torch
    assert torch.cuda.is_available()
    from vocalie_tts_b200.hift import HiFTVocoder
    sd = H.make_state_dict(0, "unit")
    return torch, sd, HiFTVocoder(sd, operand="fp16")


def _expected_reference_job(wavs, gap_ms=250, trim=True, normalize=True, target_db=-1.0):
    """The reference's order on the GPU's own waveforms: _apply_inter_chunk_gap -> PCM_16 raw file -> sf.read ->
    apply_minimal_edit (oracle restatements pinned to the reference by tests/golden)."""
    stitched = po.apply_inter_chunk_gap(wavs, sr=SR, gap_ms=gap_ms) if (gap_ms > 0 and len(wavs) > 1) else np.concatenate(wavs)
    raw = po.pcm16_encode(stitched)
    y, res = po.apply_minimal_edit_array(po.pcm16_decode(raw), SR, trim_enabled=trim, normalize_enabled=normalize, target_dbfs=target_db)
    return y.astype(np.float32), raw, res


def _expected_job(wavs, gap_ms=250, trim=True, normalize=True, target_db=-1.0):
    """numpy oracle of the opt-in per-chunk granularity applied to the GPU's own waveforms (bit-exact bar)."""
    chunks, ranges = [], []
    n = len(wavs)
    gap_on = gap_ms > 0 and n > 1
    for i, x in enumerate(wavs):
        s, e = po.trim_range_snapped(x, SR) if trim else (0, x.size)
        y = x[s:e].copy()
        if gap_on:
            if i < n - 1:
                po.fade_out(y, 240)
            if i > 0:
                po.fade_in(y, 240)
        peak = float(np.max(np.abs(y))) if y.size else 0.0
        if normalize and peak > 0:
            y = y * (float(10 ** (target_db / 20.0)) / peak)
        chunks.append(y.astype(np.float32))
        ranges.append((s, e))
    gap = np.zeros(po.ms_to_frames(SR, gap_ms) if gap_on else 0, np.float32)
    parts = []
    for i, c in enumerate(chunks):
        parts.append(c)
        if i < n - 1:
            parts.append(gap)
    return np.concatenate(parts), ranges


@pytest.mark.parametrize("gap_ms", [250, 0])
def test_pipeline_matches_oracles(setup, gap_ms):
    torch, sd, voc = setup
    from vocalie_tts_b200.pipeline import VocoderPipeline
    Ts = [30, 12, 45, 3]
    mels = [H.synth_mel(T, 11, b) for b, T in enumerate(Ts)]
    f0s = [H.synth_f0(T, 11, b) for b, T in enumerate(Ts)]
    pn = [H.synth_noise(T, 11, b) for b, T in enumerate(Ts)]
    mel, T = voc.pack_mels(mels)
    kw = dict(f0=torch.cat(f0s).cuda(), phase_vec=torch.stack([p for p, _ in pn]).cuda(),
              noise=torch.cat([n.reshape(-1) for _, n in pn]).cuda())
    raw = [w.cpu().numpy() for w in voc.inference(mels, f0=f0s, phase_vec=[p for p, _ in pn], noise=[n for _, n in pn])]
    # HiFT parity of the raw chunks
    W = H.fold_weight_norm(sd)
    for b in range(len(Ts)):
        ref = H.hift_inference(mels[b], W, f0=f0s[b], phase_vec=pn[b][0], noise=pn[b][1])
        assert H.snr_db(ref, torch.from_numpy(raw[b])) >= 60.0
    # default granularity = the reference's order (stitch raw -> PCM_16 file -> one whole-file edit)
    jp = VocoderPipeline(voc, chunk_gap_ms=gap_ms)
    jr = jp.run_device(mel, T, read_back=True, **kw)
    jwant, jraw, jmeta = _expected_reference_job(raw, gap_ms=gap_ms)
    jout = jr.audio[: jr.total_samples].cpu().numpy()
    assert np.array_equal(jr.raw[: jr.raw_samples].cpu().numpy(), jraw)
    assert jout.size == jwant.size and np.array_equal(jout.view(np.uint32), jwant.view(np.uint32))
    assert jr.edit["peak_before"] == jmeta["peak_before"] and jr.edit["gain"] == jmeta["gain"] and jr.edit["trimmed"] == jmeta["trimmed"]
    # opt-in per-chunk granularity
    pipe = VocoderPipeline(voc, chunk_gap_ms=gap_ms, granularity="chunk")
    res = pipe.run_device(mel, T, read_back=True, **kw)
    out = res.audio[: res.total_samples].cpu().numpy()
    want, ranges = _expected_job(raw, gap_ms=gap_ms)
    for i, (s, e) in enumerate(ranges):
        assert (int(res.segments[i, 0]), int(res.segments[i, 1])) == (s, e), i
    assert out.size == want.size
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_host_api_roundtrip_and_pcm16(setup):
    torch, sd, voc = setup
    from vocalie_tts_b200.pipeline import VocoderPipeline
    Ts = np.array([20, 35], np.int32)
    mel = torch.cat([H.synth_mel(int(T), 12, b).t() for b, T in enumerate(Ts)]).contiguous()
    pipe = VocoderPipeline(voc, chunk_gap_ms=250, granularity="chunk")
    a = pipe.run(mel.numpy(), Ts, seed=3)
    b = pipe.run(mel.numpy(), Ts, seed=3)
    assert a.audio.dtype == np.float32 and a.total_samples == a.audio.size
    assert np.array_equal(a.audio, b.audio)                       # deterministic for a fixed seed
    assert np.max(np.abs(a.audio)) <= 1.0
    # every chunk was normalised to -1 dBFS (fp32 rounding of x * scale)
    for i in range(len(Ts)):
        d, n = int(a.segments[i, 4]), int(a.segments[i, 5])
        assert abs(float(np.max(np.abs(a.audio[d:d + n]))) - 10 ** (-1 / 20)) < 1e-6
    pipe16 = VocoderPipeline(voc, chunk_gap_ms=250, out_pcm16=True, granularity="chunk")
    c = pipe16.run(mel.numpy(), Ts, seed=3)
    assert c.audio.dtype == np.int16
    assert np.array_equal(c.audio, po.pcm16_encode(a.audio))
    # reference order through the host API: ONE peak for the whole file, float and PCM_16 outputs agree
    j = VocoderPipeline(voc, chunk_gap_ms=250).run(mel.numpy(), Ts, seed=3)
    j16 = VocoderPipeline(voc, chunk_gap_ms=250, out_pcm16=True).run(mel.numpy(), Ts, seed=3)
    assert j.audio.dtype == np.float32 and j16.audio.dtype == np.int16 and j.total_samples == j16.total_samples
    assert np.array_equal(j16.audio, po.pcm16_encode(j.audio))
    assert abs(float(np.max(np.abs(j.audio))) - 10 ** (-1 / 20)) < 1e-6 and j.edit["normalized"]


@pytest.mark.parametrize("granularity", ["job", "chunk"])
def test_submit_collect_pipelines_jobs_with_identical_results(setup, granularity):
    """The asynchronous serving loop (two jobs in flight, device-to-host copies on a second stream) returns exactly
    what the synchronous run() returns, job by job, and refuses a third job in flight."""
    torch, sd, voc = setup
    from vocalie_tts_b200.pipeline import VocoderPipeline
    jobs = [np.array([20, 35], np.int32), np.array([12], np.int32), np.array([30, 9, 17], np.int32), np.array([25, 25], np.int32)]
    mels = [torch.cat([H.synth_mel(int(T), 40 + j, b).t() for b, T in enumerate(Ts)]).contiguous().numpy() for j, Ts in enumerate(jobs)]
    pipe = VocoderPipeline(voc, chunk_gap_ms=250, granularity=granularity)
    want = []
    for j, Ts in enumerate(jobs):
        r = pipe.run(mels[j], Ts, seed=j)
        want.append((r.audio.copy(), r.total_samples, r.segments.copy(), r.edit))
    got, pending = [], None
    for j, Ts in enumerate(jobs):
        t = pipe.submit(mels[j], Ts, seed=j)
        if pending is not None:
            r = pipe.collect(pending)
            got.append((r.audio.copy(), r.total_samples, r.segments.copy(), r.edit))
        pending = t
    with pytest.raises(RuntimeError):
        pipe.submit(mels[0], jobs[0], seed=0)
        pipe.submit(mels[0], jobs[0], seed=0)          # slot of the uncollected job
    # drain whatever is in flight, oldest first
    for t in range(pending, pipe._n_submitted):
        r = pipe.collect(t)
        if t == pending:
            got.append((r.audio.copy(), r.total_samples, r.segments.copy(), r.edit))
    assert len(got) == len(want)
    for (ga, gn, gs, ge), (wa, wn, ws, we) in zip(got, want):
        assert gn == wn and np.array_equal(ga, wa) and np.array_equal(gs, ws) and ge == we


def test_backend_drop_in(setup, tmp_path):
    torch, sd, voc = setup
    from vocalie_tts_b200 import backend as B, BackendUnavailableError
    calls = []

    def provider(text, voice_ref_path=None, lang=None, **params):
        calls.append((text, lang, params))
        T = 10 + len(text)
        return {"mel": H.synth_mel(T, 13, len(text)), "f0": H.synth_f0(T, 13, len(text))}

    B.ChatterboxB200Backend.configure(vocoder=voc, mel_provider=provider)
    try:
        assert B.ChatterboxB200Backend.is_available() and B.ChatterboxB200Backend.unavailable_reason() is None
        be = B.TTSBackend._REGISTRY["chatterbox"]()
        audio, sr, meta = be.synthesize_chunk("Bonjour le monde.", lang="fr-FR", temperature=0.7, voice="ignored")
        assert sr == SR and audio.dtype == np.float32 and audio.ndim == 1
        assert audio.size == 480 * (10 + len("Bonjour le monde."))
        assert meta["retry"] is False
        assert calls[-1][2]["temperature"] == 0.7 and "voice" not in calls[-1][2]
        out = tmp_path / "o.wav"
        m = be.synthesize("Salut.", str(out), lang="fr-FR")
        assert set(m) >= {"backend_id", "backend_lang", "out_path", "duration_s", "retry"}
        with wave.open(str(out), "rb") as w:
            assert (w.getframerate(), w.getnchannels(), w.getsampwidth()) == (SR, 1, 2)
            assert w.getnframes() == 480 * (10 + len("Salut."))
        with pytest.raises(ValueError):
            be.synthesize_chunk("   ")

        def broken(text, **kw):
            raise KeyError("boom")
        B.ChatterboxB200Backend.configure(vocoder=voc, mel_provider=broken)
        with pytest.raises(BackendUnavailableError):
            be.synthesize_chunk("x")
    finally:
        B.ChatterboxB200Backend.reset()
