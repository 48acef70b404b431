"""CPU pins of the resampler oracle: length rule of librosa.resample, the polyphase arithmetic against
scipy.signal.resample_poly with the same taps, and the filter's frequency-domain properties."""
import numpy as np
import pytest

from oracle import resample_oracle as ro


@pytest.mark.parametrize("orig,target", [(22050, 24000), (16000, 24000), (44100, 24000), (48000, 24000), (24000, 22050)])
def test_polyphase_arithmetic_matches_scipy(orig, target):
    from scipy.signal import resample_poly
    rng = np.random.default_rng(orig + target)
    for n in (1, 7, 1000, 4411):
        x = rng.standard_normal(n)
        up, down = ro.ratio(orig, target)
        got = ro.resample(x, orig, target).astype(np.float64)
        want = resample_poly(x, up, down, window=ro.design(up, down))
        assert got.size == ro.out_length(n, orig, target) == want.size, (n, got.size, want.size)
        assert np.max(np.abs(got - want)) <= 5e-7 * max(1.0, np.max(np.abs(want))), (orig, target, n)   # float32 output rounding


def test_length_rule_is_librosas_ceil():
    # int(np.ceil(n * (target / orig))) in float64 - including the cases where the product lands on an integer
    # the rule's float64 quirk is part of the contract: 22050 * (24000 / 22050) = 24000.000000000004 -> 24001 samples
    assert ro.out_length(22050, 22050, 24000) == 24001
    assert ro.out_length(44100, 44100, 24000) == int(np.ceil(44100 * (24000 / 44100)))
    assert ro.out_length(147, 22050, 24000) == 161          # 160.00000000000003 -> 161 (same quirk)
    assert ro.out_length(148, 22050, 24000) == int(np.ceil(148 * (24000 / 22050)))
    assert ro.out_length(0, 22050, 24000) == 0
    assert ro.out_length(1, 48000, 24000) == 1
    assert ro.out_length(3, 48000, 24000) == 2


def test_filter_passes_the_band_and_rejects_images():
    up, down = ro.ratio(22050, 24000)
    h = ro.design(up, down)
    H = np.abs(np.fft.rfft(h, 1 << 20))
    f = np.arange(H.size) / (1 << 20) * 2 * max(up, down)        # 1.0 = Nyquist of the lower rate
    assert np.max(np.abs(20 * np.log10(H[f <= 0.85]))) < 0.01     # flat to 0.01 dB up to 0.85 Nyquist
    assert 20 * np.log10(np.max(H[f >= 1.0])) < -110.0            # images / aliases >= 110 dB down
    # a 1 kHz tone keeps amplitude and frequency
    t = np.arange(22050) / 22050.0
    y = ro.resample(np.sin(2 * np.pi * 1000 * t), 22050, 24000)
    tt = np.arange(y.size) / 24000.0
    mid = slice(2000, -2000)
    assert np.max(np.abs(y[mid] - np.sin(2 * np.pi * 1000 * tt)[mid])) < 1e-5
