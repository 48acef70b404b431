"""One small invocation of the hot path on cuda:0 checked against the oracle
(``__graft_entry__.smoke()``): mel -> HiFT -> trim / normalise / gap stitch."""
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
    # post: per-chunk trim + normalise + gap stitch of the GPU waveforms vs the numpy oracle (bit-exact)
    pipe = VocoderPipeline(voc)
    mel, T = voc.pack_mels(mels)
    f0p = torch.cat(f0s).cuda()
    pv = torch.stack([p for p, _ in pn]).cuda()
    nz = torch.cat([n.reshape(-1) for _, n in pn]).cuda()
    res = pipe.run_device(mel, T, f0=f0p, phase_vec=pv, noise=nz, read_back=True)
    out = res.audio[: res.total_samples].cpu().numpy()
    chunks = []
    for i, x in enumerate(refs):
        s, e = po.trim_range_snapped(x, 24000)
        y = x[s:e].copy()
        if i < len(refs) - 1:
            po.fade_out(y, 240)
        if i > 0:
            po.fade_in(y, 240)
        peak = float(np.max(np.abs(y))) if y.size else 0.0
        if peak > 0:
            y = y * (float(10 ** (-1.0 / 20.0)) / peak)
        chunks.append(y.astype(np.float32))
        assert (int(res.segments[i, 0]), int(res.segments[i, 1])) == (s, e), "trim indices differ from the oracle"
    want = np.concatenate([chunks[0], np.zeros(6000, np.float32), chunks[1]])
    assert out.size == want.size, (out.size, want.size)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), "post-processing is not bit-exact"
    print(f"smoke ok: {sum(Ts) * 480} samples, {pipe.last_launches} kernel launches, operand=fp16")
