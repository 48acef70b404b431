"""HiFT vocoder oracle (PyTorch, CPU, fp32/fp64).  TEST INFRASTRUCTURE - not a product path.

PARITY UNPINNED.  The arithmetic of this half of the hot path lives in the third-party
dependency ``chatterbox-tts==0.1.6`` (reference pin: ``requirements-chatterbox.txt:1``,
``requirements-chatterbox.lock.txt:13``; call site ``tts_backends/chatterbox_impl.py:189``
``tts.generate`` -> ``S3Gen.inference`` -> ``HiFTGenerator.inference``).  That package is
not under ``/root/reference``, is not installed in the image and cannot be downloaded
(no network); the reference's own tests assert no vocoder output sample
(``tests/test_chatterbox_backend_runner.py:13-42`` are protocol mocks,
``tests/test_chatterbox_runner_venv.py:9-30`` only checks ``ok``).  This file therefore
restates the *published algorithm* of upstream
``chatterbox/models/s3gen/{hifigan.py,f0_predictor.py,s3gen.py,const.py}`` (itself
derived from CosyVoice ``hifigan/generator.py``) as instantiated by ``S3Token2Wav``:

    HiFTGenerator(sampling_rate=24000, upsample_rates=[8,5,3], upsample_kernel_sizes=[16,11,7],
                  source_resblock_kernel_sizes=[7,7,11], source_resblock_dilation_sizes=[[1,3,5]]*3,
                  in_channels=80, base_channels=512, nb_harmonics=8, nsf_alpha=0.1, nsf_sigma=0.003,
                  nsf_voiced_threshold=10, istft n_fft=16 hop=4, resblock_kernel_sizes=[3,7,11],
                  resblock_dilation_sizes=[[1,3,5]]*3, lrelu_slope=0.1, audio_limit=0.99,
                  f0_predictor=ConvRNNF0Predictor())

and anchors parity on identical (mel, F0, phase_vec, noise, weights) inputs: the CUDA
path must match this oracle to max-abs 1e-3 / SNR 60 dB with exactly ``480*T`` samples.
Randomness that upstream draws from torch's global RNG (``phase_vec``, SineGen noise)
is an explicit input here.

Differences from upstream that do not change results: functional style over a flat weight
dict with weight-norm already folded (``w = g * v / ||v||``), and per-sequence evaluation
(upstream runs batch 1).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

SR = 24000                       # upstream const.py S3GEN_SR
N_MEL = 80
BASE_CH = 512
UP_RATES = (8, 5, 3)
UP_KERNELS = (16, 11, 7)
RB_KERNELS = (3, 7, 11)
RB_DILATIONS = (1, 3, 5)
SRC_RB_KERNELS = (7, 7, 11)
N_FFT = 16
HOP = 4
NB_HARM = 8                      # 9 sine channels
SINE_AMP = 0.1
NOISE_STD = 0.003
VOICED_THR = 10.0
LRELU = 0.1
AUDIO_LIMIT = 0.99
SAMPLES_PER_FRAME = 8 * 5 * 3 * HOP  # 480
F0_CH = 512
UNIT_BRANCH_GAIN = 0.5
UNIT_POST_GAIN = 0.25


@dataclass(frozen=True)
class HiftConfig:
    """Constructor arguments of upstream ``HiFTGenerator`` that differ between its users (the module class is the same:
    Chatterbox's hifigan.py is CosyVoice's cosyvoice/hifigan/generator.py)."""
    name: str = "chatterbox_s3gen"
    sampling_rate: int = SR
    upsample_rates: tuple = UP_RATES
    upsample_kernel_sizes: tuple = UP_KERNELS
    source_resblock_kernel_sizes: tuple = SRC_RB_KERNELS
    trim_fade: bool = True            # S3Token2Wav.inference tail (Chatterbox only)

    @property
    def n_levels(self) -> int:
        return len(self.upsample_rates)

    @property
    def samples_per_frame(self) -> int:
        return int(np.prod(self.upsample_rates)) * HOP

    def source_downs(self):
        """(k, stride, pad) per level - HiFTGenerator.__init__: downsample_rates = [1] + upsample_rates[::-1][:-1],
        cumulative products reversed; u == 1 -> Conv1d(k=1) else Conv1d(k=2u, stride=u, padding=u//2)."""
        rates = [1] + list(self.upsample_rates[::-1][:-1])
        cum = list(np.cumprod(rates))[::-1]
        return [(1, 1, 0) if u == 1 else (int(2 * u), int(u), int(u // 2)) for u in cum]


CHATTERBOX = HiftConfig()
# CosyVoice-300M (v1) HiFT: cosyvoice.yaml `hift:` block - 22.05 kHz, two upsampling stages, hop 256; the default rate of the
# reference's CosyVoice runner (tts_backends/cosyvoice_runner.py:84,131)
COSYVOICE_300M = HiftConfig(name="cosyvoice_300m", sampling_rate=22050, upsample_rates=(8, 8), upsample_kernel_sizes=(16, 16),
                            source_resblock_kernel_sizes=(7, 11), trim_fade=False)
CONFIGS = {c.name: c for c in (CHATTERBOX, COSYVOICE_300M)}


# ----------------------------------------------------------------------------- weights
def _wn_names(prefix):
    return prefix + ".parametrizations.weight.original0", prefix + ".parametrizations.weight.original1"


def layer_table(cfg: "HiftConfig" = None):
    """(name, kind, C_in, C_out, k, weight_normed) for every conv, in upstream module naming."""
    cfg = cfg or CHATTERBOX
    t = []
    t.append(("conv_pre", "conv", N_MEL, BASE_CH, 7, True))
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        t.append((f"ups.{i}", "convT", BASE_CH >> i, BASE_CH >> (i + 1), k, True))
    for i, (k, s, p) in enumerate(cfg.source_downs()):
        t.append((f"source_downs.{i}", "conv", N_FFT + 2, BASE_CH >> (i + 1), k, False))
    for i in range(cfg.n_levels):
        ch = BASE_CH >> (i + 1)
        for j in range(3):
            t.append((f"source_resblocks.{i}.convs1.{j}", "conv", ch, ch, cfg.source_resblock_kernel_sizes[i], True))
            t.append((f"source_resblocks.{i}.convs2.{j}", "conv", ch, ch, cfg.source_resblock_kernel_sizes[i], True))
    for i in range(cfg.n_levels):
        ch = BASE_CH >> (i + 1)
        for kk, k in enumerate(RB_KERNELS):
            r = i * 3 + kk
            for j in range(3):
                t.append((f"resblocks.{r}.convs1.{j}", "conv", ch, ch, k, True))
                t.append((f"resblocks.{r}.convs2.{j}", "conv", ch, ch, k, True))
    t.append(("conv_post", "conv", BASE_CH >> cfg.n_levels, N_FFT + 2, 7, True))
    for i in range(5):
        t.append((f"f0_predictor.condnet.{2 * i}", "conv", N_MEL if i == 0 else F0_CH, F0_CH, 3, True))
    return t


def make_state_dict(seed: int = 0, kind: str = "init", cfg: "HiftConfig" = None) -> Dict[str, torch.Tensor]:
    """Random weights in upstream state-dict naming (weight-norm as original0=g, original1=v).

    kind="init": upstream initialisation - ``normal(0, 0.01)`` on the weight-normed convs of
      ups / resblocks / source_resblocks / conv_post (``init_weights``), PyTorch default
      (kaiming-uniform a=sqrt(5): U(+-1/sqrt(fan_in))) on conv_pre, source_downs, F0 predictor,
      l_linear, classifier; Snake alpha = 1; weight-norm g = ||v||.
    kind="unit": every conv ``normal(0, 1/sqrt(C_in*k))`` so activations stay O(1) and the
      operand rounding of the tensor-core path is actually exercised; Snake alpha ~ U(0.5, 2).
    kind="stress": "unit" pushed to what a trained checkpoint may hold - Snake alpha log-uniform in [0.05, 30];
      the trunk scaled so that activations reach |x| ~ 50 (conv_pre / source_downs x 40, conv_post / 40); and in every
      ResBlock pair a quarter of conv1's output channels carry a weight-norm gain 10^-4.5 .. 10^-1 times smaller (rows
      down to ~1e-6, fp16-subnormal) compensated by the matching input columns of conv2 - the function is (almost)
      unchanged, the operand ranges are not.
    """
    cfg = cfg or CHATTERBOX
    if kind == "stress":
        sd = make_state_dict(seed, "unit", cfg)
        g = torch.Generator().manual_seed(seed + 7919)
        for k in list(sd):
            if k.endswith(".alpha"):
                sd[k] = torch.exp(torch.empty_like(sd[k]).uniform_(math.log(0.05), math.log(30.0), generator=g))
        for k in ("conv_pre.parametrizations.weight.original0", "conv_pre.bias"):
            sd[k] = sd[k] * 40.0
        for i in range(cfg.n_levels):
            sd[f"source_downs.{i}.weight"] = sd[f"source_downs.{i}.weight"] * 40.0
            sd[f"source_downs.{i}.bias"] = sd[f"source_downs.{i}.bias"] * 40.0
        sd["conv_post.parametrizations.weight.original0"] = sd["conv_post.parametrizations.weight.original0"] / 40.0
        for k in list(sd):
            if ".convs1." in k and k.endswith("original0"):
                pre = k[: -len(".parametrizations.weight.original0")]
                c = sd[k].shape[0]
                r = torch.ones(c)
                idx = torch.rand(c, generator=g) < 0.25
                r[idx] = 10 ** torch.empty(int(idx.sum())).uniform_(-4.5, -1.0, generator=g)
                sd[k] = sd[k] * r.reshape(-1, 1, 1)
                sd[pre + ".bias"] = sd[pre + ".bias"] * r
                p2 = pre.replace("convs1", "convs2")
                v = sd[p2 + ".parametrizations.weight.original1"]
                v2 = v / r.reshape(1, -1, 1)
                gk = p2 + ".parametrizations.weight.original0"
                sd[gk] = sd[gk] * v2.flatten(1).norm(dim=1).reshape(-1, 1, 1) / v.flatten(1).norm(dim=1).reshape(-1, 1, 1)
                sd[p2 + ".parametrizations.weight.original1"] = v2
        return sd
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    for name, kind_l, cin, cout, k, wn in layer_table(cfg):
        shape = (cin, cout, k) if kind_l == "convT" else (cout, cin, k)
        fan_in = (cout if kind_l == "convT" else cin) * k  # torch's fan_in uses dim 1
        normal_init = name.startswith(("ups", "resblocks", "source_resblocks", "conv_post"))
        if kind == "unit":
            eff = cin * k / (cfg.upsample_rates[int(name.split(".")[1])] if kind_l == "convT" else 1)
            gain = 1.0
            if "resblocks" in name:
                gain = UNIT_BRANCH_GAIN      # residual branches matter but do not blow up
            elif name == "conv_post":
                gain = UNIT_POST_GAIN        # log-magnitude / phase arguments stay O(1)
            v = torch.randn(shape, generator=g) * (gain / math.sqrt(eff))
        elif normal_init:
            v = torch.randn(shape, generator=g) * 0.01
        else:
            v = uni(shape, 1.0 / math.sqrt(fan_in))
        b = uni((cout,), 1.0 / math.sqrt(fan_in))
        if wn:
            n0, n1 = _wn_names(name)
            norm = v.flatten(1).norm(dim=1).reshape(-1, 1, 1)
            if kind == "unit":
                norm = norm * (0.75 + 0.5 * torch.rand(norm.shape, generator=g))
            sd[n0] = norm
            sd[n1] = v
        else:
            sd[name + ".weight"] = v
        sd[name + ".bias"] = b
    for pre, nblk in (("resblocks", 3 * cfg.n_levels), ("source_resblocks", cfg.n_levels)):
        for r in range(nblk):
            ch = BASE_CH >> ((r // 3 if pre == "resblocks" else r) + 1)
            for j in range(3):
                for a in ("activations1", "activations2"):
                    if kind == "unit":
                        sd[f"{pre}.{r}.{a}.{j}.alpha"] = 0.5 + 1.5 * torch.rand(ch, generator=g)
                    else:
                        sd[f"{pre}.{r}.{a}.{j}.alpha"] = torch.ones(ch)
    sd["m_source.l_linear.weight"] = uni((1, NB_HARM + 1), 1.0 / 3.0)
    sd["m_source.l_linear.bias"] = uni((1,), 1.0 / 3.0)
    sd["f0_predictor.classifier.weight"] = uni((1, F0_CH), 1.0 / math.sqrt(F0_CH))
    sd["f0_predictor.classifier.bias"] = uni((1,), 1.0 / math.sqrt(F0_CH))
    if kind == "unit":
        # make the F0 head produce speech-range values (tens to hundreds of Hz)
        sd["f0_predictor.classifier.weight"] = sd["f0_predictor.classifier.weight"] * 200.0
        sd["f0_predictor.classifier.bias"] = sd["f0_predictor.classifier.bias"] + 120.0
    return sd


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``w = g * v / ||v||_2`` over all dims but 0 (torch.nn.utils.parametrizations.weight_norm,
    dim=0; for ConvTranspose1d dim 0 is C_in).  Plain ``.weight`` entries pass through."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".parametrizations.weight.original1"):
            base = k[: -len(".parametrizations.weight.original1")]
            gk = base + ".parametrizations.weight.original0"
            vv = v.double()
            norm = vv.flatten(1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
            out[base + ".weight"] = (sd[gk].double() * vv / norm).to(v.dtype)
        elif k.endswith(".parametrizations.weight.original0"):
            continue
        else:
            out[k] = v
    return out


# ----------------------------------------------------------------------------- synthetic inputs
def synth_mel(T: int, seed: int, b: int = 0) -> torch.Tensor:
    """clamp(N(-5, 2^2), ln(1e-5), 2), rounded to bf16 (the kernel's input dtype)."""
    g = torch.Generator().manual_seed(seed * 1000 + b)
    m = torch.randn(N_MEL, T, generator=g) * 2.0 - 5.0
    m = m.clamp(math.log(1e-5), 2.0)
    return m.to(torch.bfloat16).to(torch.float32)


def synth_f0(T: int, seed: int, b: int = 0) -> torch.Tensor:
    """10-frame blocks voiced w.p. 0.7; voiced value 110 + 90*(0.5+0.5*sin(2*pi*t/87 + b)) Hz."""
    g = torch.Generator().manual_seed(seed * 1000 + 500 + b)
    nblk = (T + 9) // 10
    voiced = (torch.rand(nblk, generator=g) < 0.7).repeat_interleave(10)[:T]
    t = torch.arange(T, dtype=torch.float32)
    f = 110.0 + 90.0 * (0.5 + 0.5 * torch.sin(2 * math.pi * t / 87.0 + b))
    return torch.where(voiced, f, torch.zeros_like(f)).to(torch.float32)


def synth_noise(T: int, seed: int, b: int = 0, cfg: "HiftConfig" = None):
    """phase_vec ~ U(-pi, pi) [9] with [0]=0, noise ~ N(0,1) [9, samples_per_frame * T]."""
    g = torch.Generator().manual_seed(seed * 1000 + 900 + b)
    pv = (torch.rand(NB_HARM + 1, generator=g) * 2 - 1) * math.pi
    pv[0] = 0.0
    nz = torch.randn(NB_HARM + 1, (cfg or CHATTERBOX).samples_per_frame * T, generator=g)
    return pv.to(torch.float32), nz.to(torch.float32)


# ----------------------------------------------------------------------------- modules
Quant = Optional[Callable[[torch.Tensor], torch.Tensor]]


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(x.dtype)


def fp16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float16).to(x.dtype)


def snake(x, alpha):
    """upstream hifigan.py Snake.forward: x + 1/(alpha+1e-9) * sin(alpha*x)^2, alpha [C]."""
    a = alpha.reshape(1, -1, 1)
    return x + (1.0 / (a + 1e-9)) * torch.pow(torch.sin(x * a), 2)


def _conv(x, w, b, q: Quant, **kw):
    if q is not None:
        x, w = q(x), q(w)
    return F.conv1d(x, w, b, **kw)


def resblock(x, W, prefix, k, q: Quant):
    """upstream hifigan.py ResBlock.forward (dilations 1,3,5; second conv dilation 1)."""
    for j, d in enumerate(RB_DILATIONS):
        xt = snake(x, W[f"{prefix}.activations1.{j}.alpha"])
        xt = _conv(xt, W[f"{prefix}.convs1.{j}.weight"], W[f"{prefix}.convs1.{j}.bias"], q,
                   dilation=d, padding=(k * d - d) // 2)
        xt = snake(xt, W[f"{prefix}.activations2.{j}.alpha"])
        xt = _conv(xt, W[f"{prefix}.convs2.{j}.weight"], W[f"{prefix}.convs2.{j}.bias"], q,
                   dilation=1, padding=(k - 1) // 2)
        x = xt + x
    return x


def f0_predictor(mel, W, q: Quant = None):
    """upstream f0_predictor.py ConvRNNF0Predictor.forward: 5x(conv k3 p1 + ELU), Linear, abs."""
    x = mel
    for i in range(5):
        x = _conv(x, W[f"f0_predictor.condnet.{2 * i}.weight"], W[f"f0_predictor.condnet.{2 * i}.bias"], q, padding=1)
        x = F.elu(x)
    x = x.transpose(1, 2)
    return torch.abs(F.linear(x, W["f0_predictor.classifier.weight"], W["f0_predictor.classifier.bias"]).squeeze(-1))


def sine_source(f0, W, phase_vec, noise, cfg: "HiftConfig" = None):
    """upstream SineGen.forward + SourceModuleHnNSF.forward with explicit randomness.

    f0 [B, T] (Hz) -> s [B, 1, 480T].  ``cumsum`` follows torch CPU semantics for fp32
    (accumulate in double, round each prefix to float).  phase_vec [B, 9], noise [B, 9, L].
    """
    dt = f0.dtype
    cfg = cfg or CHATTERBOX
    f0u = f0.repeat_interleave(cfg.samples_per_frame, dim=1).unsqueeze(1)      # nearest upsample [B,1,L]
    F_mat = torch.zeros(f0.size(0), NB_HARM + 1, f0u.size(-1), dtype=dt, device=f0.device)
    for i in range(NB_HARM + 1):
        F_mat[:, i:i + 1, :] = f0u * (i + 1) / cfg.sampling_rate
    theta = 2 * np.pi * (torch.cumsum(F_mat, dim=-1) % 1)
    sine = SINE_AMP * torch.sin(theta + phase_vec.reshape(f0.size(0), -1, 1).to(dt))
    uv = (f0u > VOICED_THR).to(dt)
    noise_amp = uv * NOISE_STD + (1 - uv) * SINE_AMP / 3
    sine = sine * uv + noise_amp * noise.to(dt)
    merged = torch.tanh(F.linear(sine.transpose(1, 2), W["m_source.l_linear.weight"].to(dt),
                                 W["m_source.l_linear.bias"].to(dt)))       # [B, L, 1]
    return merged.transpose(1, 2)


def stft_source(s):
    """upstream HiFTGenerator._stft: [B, L] -> real||imag [B, 18, L/4+1]."""
    if s.dtype in (torch.bfloat16, torch.float16):
        s = s.float()              # only under autocast (bench.py's stock-PyTorch GPU arm): cuFFT has no bf16
    win = torch.hann_window(N_FFT, periodic=True, dtype=s.dtype, device=s.device)
    spec = torch.stft(s, N_FFT, HOP, N_FFT, window=win, return_complex=True)
    return torch.cat([spec.real, spec.imag], dim=1)


def istft_head(mag, phase):
    """upstream HiFTGenerator._istft: clip(mag, max=100); mag*cos/sin(phase); torch.istft."""
    if mag.dtype in (torch.bfloat16, torch.float16):
        mag, phase = mag.float(), phase.float()          # autocast arm only (see stft_source)
    mag = torch.clip(mag, max=1e2)
    real = mag * torch.cos(phase)
    img = mag * torch.sin(phase)
    win = torch.hann_window(N_FFT, periodic=True, dtype=mag.dtype, device=mag.device)
    return torch.istft(torch.complex(real, img), N_FFT, HOP, N_FFT, window=win)


def trim_fade(dtype=torch.float32, sr: int = SR):
    """upstream s3gen.py S3Token2Wav.__init__: n = sr // 50 (480) zeros then (cos(linspace(pi,0,n))+1)/2."""
    n = sr // 50
    tf = torch.zeros(2 * n, dtype=dtype)
    tf[n:] = (torch.cos(torch.linspace(math.pi, 0, n, dtype=dtype)) + 1) / 2
    return tf


def decode(mel, s, W, q: Quant = None, taps: Optional[dict] = None, cfg: "HiftConfig" = None):
    """upstream HiFTGenerator.decode."""
    cfg = cfg or CHATTERBOX
    NL = cfg.n_levels
    downs = cfg.source_downs()
    s_stft = stft_source(s.squeeze(1))
    x = _conv(mel, W["conv_pre.weight"], W["conv_pre.bias"], q, padding=3)
    if taps is not None:
        taps["s_stft"] = s_stft
        taps["conv_pre"] = x
    for i in range(NL):
        x = F.leaky_relu(x, LRELU)
        xi, wi = (q(x), q(W[f"ups.{i}.weight"])) if q is not None else (x, W[f"ups.{i}.weight"])
        x = F.conv_transpose1d(xi, wi, W[f"ups.{i}.bias"], stride=cfg.upsample_rates[i],
                               padding=(cfg.upsample_kernel_sizes[i] - cfg.upsample_rates[i]) // 2)
        if i == NL - 1:
            x = F.pad(x, (1, 0), mode="reflect")
        down = downs[i][1:]
        # source_downs operate on the fp32 STFT on CUDA cores in the product (no operand rounding)
        si = F.conv1d(s_stft, W[f"source_downs.{i}.weight"], W[f"source_downs.{i}.bias"],
                      stride=down[0], padding=down[1])
        if taps is not None:
            taps[f"ups{i}"] = x
            taps[f"sd{i}"] = si
        si = resblock(si, W, f"source_resblocks.{i}", cfg.source_resblock_kernel_sizes[i], q)
        x = x + si
        if taps is not None:
            taps[f"si{i}"] = si
            taps[f"x{i}"] = x
        xs = None
        for j, k in enumerate(RB_KERNELS):
            r = resblock(x, W, f"resblocks.{i * 3 + j}", k, q)
            xs = r if xs is None else xs + r
        x = xs / 3
        if taps is not None:
            taps[f"stage{i}"] = x
    x = F.leaky_relu(x)  # default slope 0.01
    x = _conv(x, W["conv_post.weight"], W["conv_post.bias"], q, padding=3)
    if taps is not None:
        taps["conv_post"] = x
    mag = torch.exp(x[:, : N_FFT // 2 + 1, :])
    phase = torch.sin(x[:, N_FFT // 2 + 1:, :])
    y = istft_head(mag, phase)
    return torch.clamp(y, -AUDIO_LIMIT, AUDIO_LIMIT)


@torch.inference_mode()
def hift_inference(mel: torch.Tensor, W: Dict[str, torch.Tensor], *, f0: Optional[torch.Tensor] = None,
                   phase_vec: torch.Tensor, noise: torch.Tensor, dtype=torch.float32,
                   quant: Quant = None, apply_trim_fade: bool = True, taps: Optional[dict] = None,
                   cfg: "HiftConfig" = None):
    """HiFTGenerator.inference + the S3Token2Wav tail, one sequence.

    mel [80, T]; f0 [T] or None (predict); phase_vec [9]; noise [9, 480T] -> wav [480T].
    ``quant`` rounds conv operands (activations and weights) to emulate the tensor-core
    operand dtype; it is None for the oracle proper.
    """
    cfg = cfg or CHATTERBOX
    Wd = {k: v.to(dtype) for k, v in W.items()}
    mel = mel.to(dtype).unsqueeze(0)
    if f0 is None:
        f0 = f0_predictor(mel, Wd)
    else:
        f0 = f0.to(dtype).unsqueeze(0)
    s = sine_source(f0, Wd, phase_vec.unsqueeze(0).to(mel.device), noise.unsqueeze(0).to(mel.device), cfg)
    if taps is not None:
        taps["f0"] = f0
        taps["s"] = s
    y = decode(mel, s, Wd, quant, taps, cfg)
    if apply_trim_fade and cfg.trim_fade:
        tf = trim_fade(dtype, cfg.sampling_rate).to(y.device)
        n = min(tf.numel(), y.size(1))
        y[:, :n] *= tf[:n]
    return y.squeeze(0)


def snr_db(ref: torch.Tensor, test: torch.Tensor) -> float:
    ref = ref.double()
    err = (test.double() - ref)
    den = float((err * err).sum())
    num = float((ref * ref).sum())
    if den == 0.0:
        return float("inf")
    return 10.0 * math.log10(num / den) if num > 0 else float("-inf")
