"""BASELINE.json configs[3] on the GPU: a ragged batch whose mel lengths span the narration range T in {50 .. 1000}
(1-20 s chunks).  Bars: exactly 480*T samples per chunk, oracle parity (>= 60 dB, <= 1e-3) on the shortest, a middle
and the longest chunk, bit equality of every chunk with the same chunk run alone, bucketed == unbucketed, and the
job's trim indices / PCM_16 file bit-exact through VocoderPipeline against the numpy oracle run on the GPU's own
waveforms."""
import numpy as np
import pytest

from oracle import hift_oracle as H
from oracle import post_oracle as po

pytestmark = pytest.mark.gpu
SR = 24000
TS = [50, 150, 250, 350, 450, 550, 650, 750, 850, 950, 1000]      # frames (seed 1004 inputs)


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    from vocalie_tts_b200.hift import HiFTVocoder
    sd = H.make_state_dict(0, "unit")
    voc = HiFTVocoder(sd, operand="fp16")
    mels = [H.synth_mel(T, 1004, b) for b, T in enumerate(TS)]
    f0s = [H.synth_f0(T, 1004, b) for b, T in enumerate(TS)]
    pn = [H.synth_noise(T, 1004, b) for b, T in enumerate(TS)]
    wavs = [w.clone() for w in voc.inference(mels, f0=f0s, phase_vec=[p for p, _ in pn], noise=[n for _, n in pn])]
    return torch, sd, voc, mels, f0s, pn, wavs


def test_lengths_and_oracle_parity(env):
    torch, sd, voc, mels, f0s, pn, wavs = env
    for b, T in enumerate(TS):
        assert wavs[b].numel() == 480 * T
        assert bool(torch.isfinite(wavs[b]).all()) and float(wavs[b].abs().max()) <= 0.99 + 1e-7
        assert float(wavs[b][:480].abs().max()) == 0.0
    W = H.fold_weight_norm(sd)
    for b in (0, 5, len(TS) - 1):
        ref = H.hift_inference(mels[b], W, f0=f0s[b], phase_vec=pn[b][0], noise=pn[b][1])
        got = wavs[b].cpu()
        assert float((got - ref).abs().max()) <= 1e-3 and H.snr_db(ref, got) >= 60.0, (TS[b], H.snr_db(ref, got))


def test_batch_equals_solo_and_bucketed(env):
    torch, sd, voc, mels, f0s, pn, wavs = env
    for b in (0, 3, 7, len(TS) - 1):
        solo = voc.inference([mels[b]], f0=[f0s[b]], phase_vec=[pn[b][0]], noise=[pn[b][1]])[0]
        assert torch.equal(solo, wavs[b]), TS[b]
    mel, T = voc.pack_mels(mels)
    kw = dict(f0=torch.cat(f0s).cuda(), phase_vec=torch.stack([p for p, _ in pn]).cuda().contiguous(),
              noise=torch.cat([n.reshape(-1) for _, n in pn]).cuda())
    whole = voc.forward_packed(mel, T, **kw).clone()
    bucketed = voc.forward_bucketed(mel, T, max_frames=1200, **kw)       # buckets of 1-4 sequences
    assert torch.equal(whole, bucketed)
    assert torch.equal(whole, torch.cat(wavs))


@pytest.mark.parametrize("granularity", ["job", "chunk"])
def test_job_post_is_bit_exact_on_the_ragged_job(env, granularity):
    torch, sd, voc, mels, f0s, pn, wavs = env
    from vocalie_tts_b200.pipeline import VocoderPipeline
    raw = [w.cpu().numpy() for w in wavs]
    seg_off = np.concatenate([[0], np.cumsum([r.size for r in raw])]).astype(np.int64)
    pipe = VocoderPipeline(voc, chunk_gap_ms=250, out_pcm16=True, granularity=granularity)
    res = pipe.post_device(torch.cat(wavs), seg_off, read_back=True)
    got = res.audio[: res.total_samples].cpu().numpy()
    if granularity == "job":
        stitched = po.pcm16_encode(po.apply_inter_chunk_gap(raw, sr=SR, gap_ms=250))
        assert np.array_equal(res.raw[: res.raw_samples].cpu().numpy(), stitched)
        y, meta = po.apply_minimal_edit_array(po.pcm16_decode(stitched), SR, trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0)
        s, e = po.find_active_range(po.pcm16_decode(stitched), threshold=0.002, min_silence_frames=480)
        assert (res.edit["start_sample"], res.edit["end_sample"]) == (s, e)
        assert res.edit["peak_before"] == meta["peak_before"]
        assert np.array_equal(got, po.pcm16_encode(y))
    else:
        parts = []
        for i, x in enumerate(raw):
            s, e = po.trim_range_snapped(x, SR)
            assert (int(res.segments[i, 0]), int(res.segments[i, 1])) == (s, e), TS[i]
            y = x[s:e].copy()
            if i < len(raw) - 1:
                po.fade_out(y, 240)
            if i > 0:
                po.fade_in(y, 240)
            peak = float(np.max(np.abs(y)))
            parts.append((y * (float(10 ** (-1.0 / 20.0)) / peak)).astype(np.float32))
            if i < len(raw) - 1:
                parts.append(np.zeros(6000, np.float32))
        assert np.array_equal(got, po.pcm16_encode(np.concatenate(parts)))
