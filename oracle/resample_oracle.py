"""Resampler oracle (numpy float64, CPU).  TEST INFRASTRUCTURE - not a product path.

The reference converts a chunk whose engine rate differs from the pipeline's 24 kHz with
``librosa.resample(audio, orig_sr=sr, target_sr=target_sr)`` (backend/shared/tts_pipeline.py:100-111, called at
:389-390; e.g. CosyVoice's 22 050 Hz, tts_backends/cosyvoice_runner.py:84,131).  librosa's default kernel is
``res_type="soxr_hq"`` - the soxr library, which is NOT in this image (nor is librosa; no network), so the exact
impulse response cannot be reproduced here:

    PARITY UNPINNED against soxr/librosa for the sample values.

What IS pinned:
  * the output LENGTH - librosa's rule ``int(np.ceil(n * target_sr / orig_sr))`` (librosa/core/audio.py ``resample``:
    ``n_samples = int(np.ceil(y.shape[axis] * ratio))`` followed by ``util.fix_length``) - a bit-exact contract;
  * the polyphase arithmetic of the stated filter, against ``scipy.signal.resample_poly`` fed the same taps
    (tests/test_oracle_resample.py), i.e. an independent implementation of y[m] = up * sum_k x[k] h[m*down - k*up].

The filter (same on the GPU): Kaiser-windowed sinc, ``ZEROS`` zero crossings per side at the lower of the two rates,
stop-band attenuation ``ATT_DB``, transition band ending AT the lower Nyquist (no aliasing / imaging above the bar).
soxr_hq is the same class of filter (linear phase, ~0.91 pass band, >= 100 dB rejection); a fixture produced by the
real librosa can be dropped into tests/golden/ to quantify the difference where librosa exists.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

ZEROS = 64        # zero crossings per side (at the lower rate): 2 * ZEROS + 1 taps per output sample when upsampling
ATT_DB = 120.0    # stop-band attenuation of the Kaiser design


def ratio(orig_sr: int, target_sr: int) -> Tuple[int, int]:
    g = math.gcd(int(orig_sr), int(target_sr))
    return int(target_sr) // g, int(orig_sr) // g          # (up, down)


def out_length(n: int, orig_sr: int, target_sr: int) -> int:
    """librosa.resample's length rule: int(ceil(n * (target_sr / orig_sr))) with the ratio in float64."""
    return int(np.ceil(n * (float(target_sr) / orig_sr)))


def design(up: int, down: int) -> np.ndarray:
    """Prototype low-pass at the up-sampled rate (float64, odd length 2 * half + 1, DC gain 1 before the x up)."""
    r = max(up, down)
    half = ZEROS * r
    beta = 0.1102 * (ATT_DB - 8.7)
    # Kaiser transition width (rad/sample at the up-sampled rate) for this length; the stop band starts at pi / r
    dw = (ATT_DB - 8.0) / (2.285 * (2 * half))
    wc = math.pi / r - dw / 2.0                      # cutoff at the middle of the transition band
    n = np.arange(-half, half + 1, dtype=np.float64)
    h = (wc / math.pi) * np.sinc(wc / math.pi * n) * np.kaiser(2 * half + 1, beta)
    return h / h.sum()


def phase_table(up: int, down: int) -> np.ndarray:
    """H[p][j + J0] = up * h[p + j * up] (zero outside the support): y[m] = sum_j H[(m*down) % up][j + J0] * x[(m*down)//up - j]."""
    h = design(up, down)
    half = (h.size - 1) // 2
    j0 = (half + up - 1) // up
    tab = np.zeros((up, 2 * j0 + 1), dtype=np.float64)
    for p in range(up):
        for j in range(-j0, j0 + 1):
            n = p + j * up
            if -half <= n <= half:
                tab[p, j + j0] = up * h[n + half]
    return tab


def resample(x: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """The GPU kernel's definition evaluated in float64 (zero extension at both ends), float32 out."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    if orig_sr == target_sr:
        return x.astype(np.float32)
    up, down = ratio(orig_sr, target_sr)
    tab = phase_table(up, down)
    j0 = (tab.shape[1] - 1) // 2
    n_out = out_length(x.size, orig_sr, target_sr)
    xp = np.concatenate([np.zeros(j0 + 1), x, np.zeros(j0 + 1 + (n_out * down) // up)])
    m = np.arange(n_out, dtype=np.int64)
    q = m * down
    p, i0 = q % up, q // up
    y = np.zeros(n_out, dtype=np.float64)
    for j in range(-j0, j0 + 1):
        y += tab[p, j + j0] * xp[i0 - j + j0 + 1]
    return y.astype(np.float32)
