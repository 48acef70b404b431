"""One small invocation of the hot path on cuda:0 checked against the oracle
(``__graft_entry__.smoke()``): mel -> HiFT -> gap stitch -> whole-file trim / normalise (reference order)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np


def run() -> None:
    import torch
    if not torch.cuda.is_available():
        from .errors import BackendUnavailableError
        raise BackendUnavailableError("smoke() needs cuda:0")
    torch.cuda.set_device(0)
    root = Path(__file__).resolve().parent.parent
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    from oracle import hift_oracle as H          # checker only
    from oracle import post_oracle as po
    from .hift import HiFTVocoder
    from .pipeline import VocoderPipeline

    Ts = [40, 25]
    sd = H.make_state_dict(0, "unit")
    W = H.fold_weight_norm(sd)
    mels = [H.synth_mel(T, 1, b) for b, T in enumerate(Ts)]
    f0s = [H.synth_f0(T, 1, b) for b, T in enumerate(Ts)]
    pn = [H.synth_noise(T, 1, b) for b, T in enumerate(Ts)]
    voc = HiFTVocoder(sd, operand="fp16")
    wavs = voc.inference(mels, f0=f0s, phase_vec=[p for p, _ in pn], noise=[n for _, n in pn])
    refs = []
    for b, T in enumerate(Ts):
        ref = H.hift_inference(mels[b], W, f0=f0s[b], phase_vec=pn[b][0], noise=pn[b][1])
        got = wavs[b].cpu()
        assert got.numel() == 480 * T, (got.numel(), T)
        err = float((got - ref).abs().max())
        snr = H.snr_db(ref, got)
        assert err <= 1e-3 and snr >= 60.0, f"HiFT parity failed: max-abs {err:.3g}, SNR {snr:.1f} dB"
        refs.append(got.numpy())
    # post in the reference's order: gap stitch of the raw chunks -> PCM_16 raw file -> one whole-file
    # apply_minimal_edit (trim, ONE peak, clip), GPU waveforms vs the numpy oracle (bit-exact)
    pipe = VocoderPipeline(voc)
    mel, T = voc.pack_mels(mels)
    f0p = torch.cat(f0s).cuda()
    pv = torch.stack([p for p, _ in pn]).cuda()
    nz = torch.cat([n.reshape(-1) for _, n in pn]).cuda()
    res = pipe.run_device(mel, T, f0=f0p, phase_vec=pv, noise=nz, read_back=True)
    out = res.audio[: res.total_samples].cpu().numpy()
    raw = po.pcm16_encode(po.apply_inter_chunk_gap(refs, sr=24000, gap_ms=250))
    assert np.array_equal(res.raw[: res.raw_samples].cpu().numpy(), raw), "stitched raw file is not bit-exact"
    want, meta = po.apply_minimal_edit_array(po.pcm16_decode(raw), 24000, trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0)
    assert out.size == want.size, (out.size, want.size)
    assert np.array_equal(out.view(np.uint32), want.astype(np.float32).view(np.uint32)), "post-processing is not bit-exact"
    assert res.edit["peak_before"] == meta["peak_before"] and res.edit["gain"] == meta["gain"]
    print(f"smoke ok: {sum(Ts) * 480} samples, {pipe.last_launches} kernel launches, operand=fp16")
