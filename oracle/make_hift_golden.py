"""Pin the HiFT oracle to the REAL upstream modules (chatterbox-tts==0.1.6) - run where they are importable.

TEST INFRASTRUCTURE.  The build container and the GPU box have no network and no ``chatterbox`` package
(reference pin: requirements-chatterbox.txt:1; call site tts_backends/chatterbox_impl.py:189), so
``oracle/hift_oracle.py`` is a restatement whose parity is UNPINNED.  This script closes that gap for anyone
who has the package:

    pip install chatterbox-tts==0.1.6
    python oracle/make_hift_golden.py            # writes tests/golden/hift_upstream.npz (~1 MB)

It builds upstream's own ``HiFTGenerator`` exactly as ``S3Token2Wav.__init__`` does (s3gen.py), loads the seeded
synthetic state dict of ``hift_oracle.make_state_dict`` into it (same parameter names, so only the SEED travels,
not 80 MB of weights), runs ``HiFTGenerator.inference(speech_feat=mel)`` followed by the ``trim_fade`` tail of
``S3Token2Wav.inference``, and records the randomness upstream drew from torch's global RNG inside
``SineGen.forward`` (``Uniform.sample`` -> phase_vec, ``randn_like`` -> noise) so the oracle and the CUDA path can
be fed identical (mel, phase_vec, noise, weights).

``tests/test_hift_upstream_golden.py`` then (a) pins the oracle: oracle(mel, ...) == upstream wav to fp32 rounding,
and (b) holds the CUDA path to the north-star bar against the UPSTREAM waveform.  Both skip while the fixture is
absent, and say so.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden" / "hift_upstream.npz"
CASES = [  # (weights kind, weights seed, mel frames, input seed, injected F0?)
    ("unit", 0, 24, 101, True),
    ("unit", 0, 16, 102, False),     # F0 from upstream's ConvRNNF0Predictor
    ("init", 0, 24, 103, True),
]


def build_upstream():
    """HiFTGenerator as instantiated by upstream S3Token2Wav.__init__."""
    from chatterbox.models.s3gen.hifigan import HiFTGenerator              # noqa: E402  (absent offline)
    from chatterbox.models.s3gen.f0_predictor import ConvRNNF0Predictor    # noqa: E402
    return HiFTGenerator(sampling_rate=24000, upsample_rates=[8, 5, 3], upsample_kernel_sizes=[16, 11, 7],
                         source_resblock_kernel_sizes=[7, 7, 11], source_resblock_dilation_sizes=[[1, 3, 5]] * 3,
                         f0_predictor=ConvRNNF0Predictor())


class _RecordRandomness:
    """Wraps the two RNG draws of upstream SineGen.forward so their values become explicit inputs."""

    def __init__(self):
        import torch
        self.torch = torch
        self.uniform, self.normal = [], []

    def __enter__(self):
        torch = self.torch
        self._sample = torch.distributions.Uniform.sample
        self._randn_like = torch.randn_like
        rec = self

        def sample(dist, sample_shape=torch.Size()):
            v = rec._sample(dist, sample_shape)
            rec.uniform.append(v.detach().clone())
            return v

        def randn_like(t, *a, **k):
            v = rec._randn_like(t, *a, **k)
            rec.normal.append(v.detach().clone())
            return v

        torch.distributions.Uniform.sample = sample
        torch.randn_like = randn_like
        return self

    def __exit__(self, *exc):
        self.torch.distributions.Uniform.sample = self._sample
        self.torch.randn_like = self._randn_like


def main():
    import torch
    sys.path.insert(0, str(ROOT))
    from oracle import hift_oracle as H
    try:
        gen = build_upstream().eval()
    except ImportError as exc:
        raise SystemExit(f"chatterbox-tts is not importable here ({exc}); install chatterbox-tts==0.1.6 first - "
                         "the HiFT oracle stays 'parity unpinned' until this script has run")
    g = {}
    meta = []
    for ci, (kind, wseed, T, seed, inject_f0) in enumerate(CASES):
        sd = H.make_state_dict(wseed, kind)
        missing, unexpected = gen.load_state_dict(sd, strict=False)
        # the synthetic dict must cover every parameter of the real module - otherwise the restated layer table is wrong
        assert not [k for k in missing if not k.endswith("num_batches_tracked")], f"oracle state dict misses upstream parameters: {missing}"
        assert not unexpected, f"oracle state dict has names upstream does not know: {unexpected}"
        mel = H.synth_mel(T, seed, 0)
        f0 = H.synth_f0(T, seed, 0) if inject_f0 else None
        torch.manual_seed(seed)
        with torch.inference_mode(), _RecordRandomness() as rec:
            if inject_f0:
                # HiFTGenerator.inference with the F0 predictor bypassed: same body, F0 given (north_star: parity on
                # identical mel / F0 inputs)
                s = gen.f0_upsamp(f0[None, None, :]).transpose(1, 2)
                s, _, _ = gen.m_source(s)
                wav = gen.decode(x=mel[None], s=s.transpose(1, 2))
            else:
                wav, _ = gen.inference(speech_feat=mel[None], cache_source=torch.zeros(1, 1, 0))
        wav = wav.reshape(-1).clone()
        # S3Token2Wav.inference tail (s3gen.py): output_wavs[:, :len(trim_fade)] *= trim_fade
        n = 24000 // 50
        tf = torch.zeros(2 * n)
        tf[n:] = (torch.cos(torch.linspace(torch.pi, 0, n)) + 1) / 2
        m = min(tf.numel(), wav.numel())
        wav[:m] *= tf[:m]
        assert len(rec.uniform) >= 1 and len(rec.normal) >= 1, "upstream SineGen drew its randomness differently than restated"
        pv = rec.uniform[0].reshape(-1)[:9].clone()
        pv[0] = 0.0
        noise = rec.normal[0].reshape(9, -1) if rec.normal[0].shape[1] == 9 else rec.normal[0].reshape(-1, 9).t()
        g[f"mel_{ci}"] = mel.numpy()
        if inject_f0:
            g[f"f0_{ci}"] = f0.numpy()
        g[f"pv_{ci}"] = pv.numpy().astype(np.float32)
        g[f"noise_{ci}"] = noise.contiguous().numpy().astype(np.float32)
        g[f"wav_{ci}"] = wav.numpy().astype(np.float32)
        meta.append({"kind": kind, "weights_seed": wseed, "T": T, "seed": seed, "inject_f0": inject_f0})
        print(f"case {ci}: {kind} T={T} -> {wav.numel()} samples")
    import chatterbox
    g["meta"] = np.frombuffer(json.dumps({"cases": meta, "chatterbox": getattr(chatterbox, "__version__", "unknown"),
                                          "torch": torch.__version__}).encode(), np.uint8)
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
