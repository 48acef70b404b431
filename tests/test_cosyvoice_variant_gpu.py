"""Second consumer of the same kernels: the CosyVoice-300M instantiation of HiFTGenerator - 22.05 kHz,
two upsampling stages (8, 8), kernels (16, 16), source ResBlock kernels (7, 11), hop 256, no trim_fade tail - behind the
same C ABI (vt_hift_create_ex).  Same bars as the Chatterbox instantiation, against the same oracle run with that
configuration: exactly 256*T samples, <= 1e-3 / >= 60 dB on the tensor-core path, tap-by-tap agreement on the exact path."""
import numpy as np
import pytest

from oracle import hift_oracle as H

pytestmark = pytest.mark.gpu
CFG = H.COSYVOICE_300M


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    sd = H.make_state_dict(0, "unit", CFG)
    Ts = [37, 50, 8, 1, 130]
    mels = [H.synth_mel(T, 23, b) for b, T in enumerate(Ts)]
    f0s = [H.synth_f0(T, 23, b) for b, T in enumerate(Ts)]
    pn = [H.synth_noise(T, 23, b, CFG) for b, T in enumerate(Ts)]
    return torch, sd, Ts, mels, f0s, pn


def test_configuration_is_reported_and_validated(env):
    torch, sd, *_ = env
    from vocalie_tts_b200.hift import HiFTVocoder, algorithmic_flops_per_frame
    voc = HiFTVocoder(sd, operand="fp32", config="cosyvoice_300m")
    assert (voc.samples_per_frame, voc.sr) == (256, 22050) == (CFG.samples_per_frame, CFG.sampling_rate)
    assert algorithmic_flops_per_frame(config="cosyvoice_300m") < algorithmic_flops_per_frame()
    with pytest.raises(ValueError):
        HiFTVocoder(sd, config="no_such_generator")
    from vocalie_tts_b200 import BackendUnavailableError
    with pytest.raises(BackendUnavailableError):            # Chatterbox weights do not fit the CosyVoice layer table
        HiFTVocoder(H.make_state_dict(0, "unit"), operand="fp32", config="cosyvoice_300m")


@pytest.mark.parametrize("operand", ["fp32", "fp16"])
def test_waveform_and_taps_match_the_oracle(env, operand):
    torch, sd, Ts, mels, f0s, pn = env
    from vocalie_tts_b200.hift import HiFTVocoder
    voc = HiFTVocoder(sd, operand=operand, config="cosyvoice_300m")
    wavs = voc.inference(mels, f0=f0s, phase_vec=[p for p, _ in pn], noise=[n for _, n in pn])
    W = H.fold_weight_norm(sd)
    tap_ch = {"conv_pre": 512, "ups0": 256, "x0": 256, "stage0": 256, "ups1": 128, "x1": 128, "stage1": 128, "conv_post": 18}
    for b, T in enumerate(Ts):
        taps = {}
        ref = H.hift_inference(mels[b], W, f0=f0s[b], phase_vec=pn[b][0], noise=pn[b][1], cfg=CFG, taps=taps)
        got = wavs[b].cpu()
        assert got.numel() == 256 * T == ref.numel()
        for name, ch in tap_ch.items():
            g = voc.read_tap(name, b, ch).cpu()
            w = taps[name][0].t().contiguous()
            assert g.shape == w.shape, (name, g.shape, w.shape)
            rel = float((g.double() - w.double()).abs().max()) / (float(w.abs().max()) + 1e-12)
            assert rel < (1e-3 if operand == "fp32" else 3e-3), (operand, name, b, rel)
        err, snr = float((got - ref).abs().max()), H.snr_db(ref, got)
        if operand == "fp32":
            assert err <= 1e-4 and snr >= 80.0, (b, err, snr)
        else:
            assert err <= 1e-3 and snr >= 60.0, (b, err, snr)
        assert float(got.abs().max()) <= 0.99 + 1e-7
    with pytest.raises(Exception):
        voc.read_tap("stage2", 0, 64)                       # this generator has two levels


def test_job_pipeline_runs_at_the_generator_rate(env):
    torch, sd, Ts, mels, f0s, pn = env
    from oracle import post_oracle as po
    from vocalie_tts_b200.hift import HiFTVocoder
    from vocalie_tts_b200.pipeline import VocoderPipeline
    voc = HiFTVocoder(sd, operand="fp16", config="cosyvoice_300m")
    pipe = VocoderPipeline(voc, chunk_gap_ms=250, out_pcm16=True)
    assert pipe.sr == 22050 and pipe.spf == 256
    mel, T = voc.pack_mels(mels[:3])
    kw = dict(f0=torch.cat(f0s[:3]).cuda(), phase_vec=torch.stack([p for p, _ in pn[:3]]).cuda().contiguous(),
              noise=torch.cat([n.reshape(-1) for _, n in pn[:3]]).cuda())
    raw = [w.cpu().numpy() for w in voc.inference(mels[:3], f0=f0s[:3], phase_vec=[p for p, _ in pn[:3]], noise=[n for _, n in pn[:3]])]
    res = pipe.run_device(mel, T, read_back=True, **kw)
    stitched = po.pcm16_encode(po.apply_inter_chunk_gap(raw, sr=22050, gap_ms=250))      # gap = int(22050 * 0.25) = 5512 samples
    assert np.array_equal(res.raw[: res.raw_samples].cpu().numpy(), stitched)
    y, meta = po.apply_minimal_edit_array(po.pcm16_decode(stitched), 22050, trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0)
    assert np.array_equal(res.audio[: res.total_samples].cpu().numpy(), po.pcm16_encode(y))
    # and the 22.05 kHz job resampled to the pipeline's 24 kHz keeps librosa's length rule (tts_pipeline.py:389-390)
    from vocalie_tts_b200 import post
    up = post._resample_audio(raw[0], 22050, 24000)
    assert up.size == int(np.ceil(raw[0].size * (24000 / 22050)))
