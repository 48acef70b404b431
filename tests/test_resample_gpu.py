"""GPU tests of the polyphase resampler behind ``_resample_audio`` (reference backend/shared/tts_pipeline.py:100-111).
Sample counts follow librosa's rule exactly; sample values are held to the float64 oracle of the SAME stated filter
(oracle/resample_oracle.py, itself pinned to scipy.signal.resample_poly) - the soxr kernel librosa would use is absent
here, so values are not bit-pinned to the reference (parity unpinned, said so in DESIGN.md)."""
import numpy as np
import pytest

from oracle import resample_oracle as ro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def post():
    import torch
    assert torch.cuda.is_available()
    from vocalie_tts_b200 import post
    return post


@pytest.mark.parametrize("orig,target", [(22050, 24000), (16000, 24000), (44100, 24000), (48000, 24000), (24000, 22050)])
def test_matches_the_float64_oracle(post, orig, target):
    rng = np.random.default_rng(orig)
    for n in (1, 5, 147, 4411, orig, 3 * orig + 17):
        x = (rng.standard_normal(n) * 0.3).astype(np.float32)
        got = post._resample_audio(x, orig, target)
        want = ro.resample(x, orig, target)
        assert got.dtype == np.float32 and got.size == want.size == ro.out_length(n, orig, target), (n, got.size)
        assert np.max(np.abs(got - want)) <= 3e-6, (orig, target, n, float(np.max(np.abs(got - want))))


def test_identity_empty_and_multichannel(post):
    x = np.arange(10, dtype=np.float32)
    assert post._resample_audio(x, 24000, 24000) is x                          # tts_pipeline.py:101-102
    assert post._resample_audio(np.zeros(0, np.float32), 22050, 24000).size == 0
    rng = np.random.default_rng(3)
    st = (rng.standard_normal((5000, 2)) * 0.2).astype(np.float32)
    y = post._resample_audio(st, 22050, 24000)
    assert y.shape == (ro.out_length(5000, 22050, 24000), 2)
    for c in range(2):
        assert np.max(np.abs(y[:, c] - ro.resample(st[:, c], 22050, 24000))) <= 3e-6


def test_batched_segments_and_tone(post):
    import torch
    rng = np.random.default_rng(9)
    lens = [1000, 1, 22050, 333]
    segs = [(rng.standard_normal(n) * 0.3).astype(np.float32) for n in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out, ooff = post.resample_device(torch.from_numpy(np.concatenate(segs)).cuda(), off, 22050, 24000)
    out = out.cpu().numpy()
    for i, s in enumerate(segs):
        want = ro.resample(s, 22050, 24000)
        assert ooff[i + 1] - ooff[i] == want.size
        assert np.max(np.abs(out[ooff[i]:ooff[i + 1]] - want)) <= 3e-6, i
    # one second at 22 050 Hz -> 24 001 samples (librosa's float64 ceil quirk), a 1 kHz tone stays a 1 kHz tone
    t = np.arange(22050) / 22050.0
    y = post._resample_audio(np.sin(2 * np.pi * 1000 * t).astype(np.float32), 22050, 24000)
    assert y.size == 24001
    tt = np.arange(y.size) / 24000.0
    assert np.max(np.abs(y[2000:-2000] - np.sin(2 * np.pi * 1000 * tt)[2000:-2000])) < 2e-5
