"""GPU parity of the HiFT vocoder (through the C ABI) against the torch fp32 oracle on identical
(mel, F0, phase_vec, noise, weights).  Bars (BASELINE.json north_star): exactly 480*T samples,
max-abs error <= 1e-3 and SNR >= 60 dB for the waveform; intermediates are checked layer by layer
with a relative tolerance so a wrong layer is named, not just a wrong waveform."""
import math

import numpy as np
import pytest

from oracle import hift_oracle as H

pytestmark = pytest.mark.gpu

MAX_ABS = 1e-3      # north_star tolerance
MIN_SNR_DB = 60.0


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available(), "GPU tests need a CUDA device"
    t.set_num_threads(max(1, min(16, t.get_num_threads())))
    return t


@pytest.fixture(scope="module")
def weights():
    return {kind: H.make_state_dict(0, kind) for kind in ("init", "unit")}


@pytest.fixture(scope="module")
def vocoders(torch, weights):
    from vocalie_tts_b200.hift import HiFTVocoder
    cache = {}

    def get(kind, operand):
        key = (kind, operand)
        if key not in cache:
            cache[key] = HiFTVocoder(weights[kind], operand=operand)
        return cache[key]
    return get


def _inputs(torch, Ts, seed):
    mels = [H.synth_mel(T, seed, b) for b, T in enumerate(Ts)]
    f0s = [H.synth_f0(T, seed, b) for b, T in enumerate(Ts)]
    pvs, nzs = zip(*[H.synth_noise(T, seed, b) for b, T in enumerate(Ts)])
    return mels, f0s, list(pvs), list(nzs)


def _oracle(torch, W, mel, f0, pv, nz, taps=None, dtype=None):
    return H.hift_inference(mel, W, f0=f0, phase_vec=pv, noise=nz, taps=taps, dtype=dtype or torch.float32)


def _rel_err(ref, got):
    ref = ref.double()
    got = got.double()
    scale = float(ref.abs().max()) + 1e-12
    return float((ref - got).abs().max()) / scale


TAP_CH = {"s": 1, "s_stft": 18, "conv_pre": 512, "ups0": 256, "x0": 256, "stage0": 256, "ups1": 128, "x1": 128,
          "stage1": 128, "ups2": 64, "x2": 64, "stage2": 64, "conv_post": 18}


@pytest.mark.parametrize("kind", ["init", "unit"])
def test_fp32_path_matches_oracle_layer_by_layer(torch, weights, vocoders, kind):
    Ts = [37, 50, 8, 1]
    mels, f0s, pvs, nzs = _inputs(torch, Ts, seed=3)
    voc = vocoders(kind, "fp32")
    wavs = voc.inference(mels, f0=f0s, phase_vec=pvs, noise=nzs)
    W = H.fold_weight_norm(weights[kind])
    for b, T in enumerate(Ts):
        taps = {}
        ref = _oracle(torch, W, mels[b], f0s[b], pvs[b], nzs[b], taps)
        for name, ch in TAP_CH.items():
            got = voc.read_tap(name, b, ch).cpu()
            want = taps[name][0].t().contiguous() if name != "s" else taps["s"][0].t().contiguous()
            assert got.shape == want.shape, (name, b, got.shape, want.shape)
            tol = 2e-3 if name in ("s", "s_stft") else 1e-3
            # the source can differ on isolated samples (phase rounding boundary): judge s by RMS too
            if name == "s":
                rms = float((got - want).double().pow(2).mean().sqrt())
                assert rms < 1e-4, (name, b, rms)
            else:
                assert _rel_err(want, got) < tol, (kind, name, b, _rel_err(want, got))
        got = wavs[b].cpu()
        assert got.numel() == 480 * T == ref.numel()
        assert float((got - ref).abs().max()) <= 1e-4, (kind, b, float((got - ref).abs().max()))
        assert H.snr_db(ref, got) >= 80.0, (kind, b, H.snr_db(ref, got))


def test_tensor_core_path_layer_by_layer(torch, weights, vocoders):
    """The tcgen05 path (K-blocked layers, fused ResBlock pairs at C = 64 / 128, activation-resident convs at
    C = 256) against the oracle, tap by tap: fp16 operands bound every intermediate to ~1e-3 of its range.
    Lengths cross the fused kernel's 246/250/254-step tiles and include sequences shorter than one halo."""
    Ts = [37, 50, 8, 1, 3]
    mels, f0s, pvs, nzs = _inputs(torch, Ts, seed=13)
    voc = vocoders("unit", "fp16")
    wavs = voc.inference(mels, f0=f0s, phase_vec=pvs, noise=nzs)
    W = H.fold_weight_norm(weights["unit"])
    for b, T in enumerate(Ts):
        taps = {}
        ref = _oracle(torch, W, mels[b], f0s[b], pvs[b], nzs[b], taps)
        for name in ("conv_pre", "ups0", "x0", "stage0", "ups1", "x1", "stage1", "ups2", "x2", "stage2", "conv_post"):
            got = voc.read_tap(name, b, TAP_CH[name]).cpu()
            want = taps[name][0].t().contiguous()
            assert got.shape == want.shape, (name, b, got.shape, want.shape)
            assert torch.isfinite(got).all(), (name, b)
            assert _rel_err(want, got) < 3e-3, (name, b, _rel_err(want, got))
        got = wavs[b].cpu()
        assert got.numel() == 480 * T == ref.numel()
        assert float((got - ref).abs().max()) <= MAX_ABS and H.snr_db(ref, got) >= MIN_SNR_DB, (b, H.snr_db(ref, got))


def test_tensor_core_path_source_taps(torch, weights, vocoders):
    """The source half on the tensor-core path: s (SineGen + tanh(Linear), fp64 phase prefix) and its 16-point STFT,
    which this path keeps only as fp16 operand rows of the source_down GEMMs."""
    Ts = [37, 50, 8]
    mels, f0s, pvs, nzs = _inputs(torch, Ts, seed=19)
    voc = vocoders("unit", "fp16")
    voc.inference(mels, f0=f0s, phase_vec=pvs, noise=nzs)
    W = H.fold_weight_norm(weights["unit"])
    for b, T in enumerate(Ts):
        taps = {}
        _oracle(torch, W, mels[b], f0s[b], pvs[b], nzs[b], taps)
        got = voc.read_tap("s", b, 1).cpu()
        want = taps["s"][0].t().contiguous()
        assert got.shape == want.shape
        assert float((got - want).double().pow(2).mean().sqrt()) < 1e-4, b      # isolated phase-boundary samples may differ
        got = voc.read_tap("s_stft", b, 18).cpu()
        want = taps["s_stft"][0].t().contiguous()
        assert got.shape == want.shape, (got.shape, want.shape)
        assert _rel_err(want, got) < 2e-3, (b, _rel_err(want, got))             # fp16 rows: 2^-11 relative per element


def test_stress_weights_meet_the_parity_bar(torch, vocoders):
    """Operand ranges a trained checkpoint may hold (oracle kind="stress"): Snake alpha in [0.05, 30], activations up to
    |x| ~ 800, weight rows down to ~1e-6 (fp16-subnormal) feeding columns up to ~1e3.  Plain fp16 weight images reach
    only ~53 dB on this set (oracle emulation); the per-row power-of-two scales folded into the packed weights and
    undone in the epilogue FMAs restore the bar."""
    from vocalie_tts_b200.hift import HiFTVocoder
    sd = H.make_state_dict(0, "stress")
    W = H.fold_weight_norm(sd)
    voc = HiFTVocoder(sd, operand="fp16")
    Ts = [60, 17, 300]
    mels, f0s, pvs, nzs = _inputs(torch, Ts, seed=31)
    wavs = voc.inference(mels, f0=f0s, phase_vec=pvs, noise=nzs)
    for b, T in enumerate(Ts):
        ref = _oracle(torch, W, mels[b], f0s[b], pvs[b], nzs[b])
        got = wavs[b].cpu()
        assert got.numel() == 480 * T and bool(torch.isfinite(got).all())
        err, snr = float((got - ref).abs().max()), H.snr_db(ref, got)
        assert err <= MAX_ABS and snr >= MIN_SNR_DB, (T, err, snr)
    # the exact CUDA-core path on the same weights (no operand rounding at all)
    voc32 = HiFTVocoder(sd, operand="fp32")
    w32 = voc32.inference(mels[:2], f0=f0s[:2], phase_vec=pvs[:2], noise=nzs[:2])
    for b in range(2):
        ref = _oracle(torch, W, mels[b], f0s[b], pvs[b], nzs[b])
        assert H.snr_db(ref, w32[b].cpu()) >= 80.0


def test_fused_pairs_agree_with_unfused_convs(torch, weights, monkeypatch):
    """VT_NO_FUSE=1 runs every ResBlock conv as its own launch (operand copies in HBM); the fused pair kernel
    must agree with it well inside the parity bar (same arithmetic, but the fp32 residual add order differs, which
    moves individual fp16 operand roundings downstream)."""
    from vocalie_tts_b200.hift import HiFTVocoder
    Ts = [123, 40]
    mels, f0s, pvs, nzs = _inputs(torch, Ts, seed=17)
    fused = HiFTVocoder(weights["unit"], operand="fp16")
    monkeypatch.setenv("VT_NO_FUSE", "1")
    plain = HiFTVocoder(weights["unit"], operand="fp16")
    monkeypatch.delenv("VT_NO_FUSE")
    a = [w.clone().cpu() for w in fused.inference(mels, f0=f0s, phase_vec=pvs, noise=nzs)]
    n_fused = fused.last_launches
    b = [w.clone().cpu() for w in plain.inference(mels, f0=f0s, phase_vec=pvs, noise=nzs)]
    assert n_fused < plain.last_launches
    for x, y in zip(a, b):
        assert H.snr_db(y, x) >= 66.0, H.snr_db(y, x)


@pytest.mark.parametrize("kind", ["init", "unit"])
@pytest.mark.parametrize("operand", ["fp32", "fp16"])
def test_waveform_parity_cfg1(torch, weights, vocoders, kind, operand):
    """configs[0]: a single ~5 s chunk (T=250) -> 120 000 samples."""
    T = 250
    mels, f0s, pvs, nzs = _inputs(torch, [T], seed=1)
    voc = vocoders(kind, operand)
    got = voc.inference(mels, f0=f0s, phase_vec=pvs, noise=nzs)[0].cpu()
    ref = _oracle(torch, H.fold_weight_norm(weights[kind]), mels[0], f0s[0], pvs[0], nzs[0])
    assert got.numel() == ref.numel() == 120000
    err = float((got - ref).abs().max())
    snr = H.snr_db(ref, got)
    assert err <= MAX_ABS, (kind, operand, err)
    assert snr >= MIN_SNR_DB, (kind, operand, snr)
    assert float(got.abs().max()) <= 0.99 + 1e-7
    assert float(got[:480].abs().max()) == 0.0          # trim_fade zeroes the first 20 ms


@pytest.mark.parametrize("operand", ["fp32", "fp16"])
def test_ragged_batch_equals_single_sequences(torch, weights, vocoders, operand):
    """Batching must not change results: every sequence of a ragged batch equals the same sequence
    run alone (edge masking / zero padding per sequence)."""
    Ts = [64, 5, 129, 33, 2]
    mels, f0s, pvs, nzs = _inputs(torch, Ts, seed=5)
    voc = vocoders("unit", operand)
    batch = [w.clone() for w in voc.inference(mels, f0=f0s, phase_vec=pvs, noise=nzs)]
    for b in range(len(Ts)):
        solo = voc.inference([mels[b]], f0=[f0s[b]], phase_vec=[pvs[b]], noise=[nzs[b]])[0]
        assert torch.equal(batch[b], solo), (operand, b, float((batch[b] - solo).abs().max()))


def test_f0_predictor_matches_oracle(torch, weights, vocoders):
    Ts = [60, 17]
    mels, _, pvs, nzs = _inputs(torch, Ts, seed=7)
    voc = vocoders("unit", "fp32")
    wavs = voc.inference(mels, phase_vec=pvs, noise=nzs)        # no F0 given -> ConvRNNF0Predictor runs
    W = H.fold_weight_norm(weights["unit"])
    for b, T in enumerate(Ts):
        want = H.f0_predictor(mels[b].unsqueeze(0), W)[0]
        got = voc.read_tap("f0", b, 1).cpu().reshape(-1)
        assert _rel_err(want, got) < 1e-4, (b, _rel_err(want, got))
        ref = _oracle(torch, W, mels[b], None, pvs[b], nzs[b])
        assert wavs[b].numel() == 480 * T
        assert H.snr_db(ref, wavs[b].cpu()) >= 60.0


def test_f0_predictor_tensor_core_split_precision(torch, weights, vocoders):
    """In the tensor-core modes the F0 trunk runs on tcgen05 with two-term fp16 operands (x = hi + lo,
    three products): F0 feeds a phase integral over the whole chunk, so it keeps ~fp32 accuracy while the
    rest of the vocoder runs on single fp16 operands.  Lengths cross the 256-row tile boundary."""
    Ts = [300, 17, 256]
    mels, _, pvs, nzs = _inputs(torch, Ts, seed=11)
    voc = vocoders("unit", "fp16")
    wavs = voc.inference(mels, phase_vec=pvs, noise=nzs)
    W = H.fold_weight_norm(weights["unit"])
    for b, T in enumerate(Ts):
        want = H.f0_predictor(mels[b].unsqueeze(0), W)[0]
        got = voc.read_tap("f0", b, 1).cpu().reshape(-1)
        assert got.numel() == T and wavs[b].numel() == 480 * T
        # measured 1.7e-5 .. 2.2e-5 of the F0 range on these inputs (3.6e-5 with the upstream-init weights): the three
        # fp16 x fp16 products are exact, what remains is the tensor core's fp32 accumulation over K = 3 x 1536 terms.
        # Plain fp16 operands give ~1e-3.
        assert _rel_err(want, got) < 5e-5, (b, _rel_err(want, got))


def test_internal_noise_mode_is_deterministic_and_bounded(torch, vocoders):
    Ts = [40, 21]
    mels, f0s, _, _ = _inputs(torch, Ts, seed=9)
    voc = vocoders("unit", "fp16")
    a = [w.clone() for w in voc.inference(mels, f0=f0s, seed=1234)]
    b = [w.clone() for w in voc.inference(mels, f0=f0s, seed=1234)]
    c = [w.clone() for w in voc.inference(mels, f0=f0s, seed=99)]
    for x, y, z, T in zip(a, b, c, Ts):
        assert x.numel() == 480 * T
        assert torch.equal(x, y)
        assert not torch.equal(x, z)
        assert bool(torch.isfinite(x).all()) and float(x.abs().max()) <= 0.99 + 1e-7
    # the in-kernel generator must look like N(0,1) noise through the unvoiced branch:
    # with F0 = 0 everywhere, s = tanh(w . (0.1/3 * z) + b) has the right spread
    zeros = [torch.zeros(T) for T in Ts]
    voc.inference(mels, f0=zeros, seed=5)
    s = voc.read_tap("s", 0, 1).cpu().reshape(-1)
    assert 0.0 < float(s.std()) < 0.2 and abs(float(s.mean())) < 0.5


def test_bad_arguments_raise_backend_error(torch, weights):
    from vocalie_tts_b200 import BackendUnavailableError
    from vocalie_tts_b200.hift import HiFTVocoder
    sd = dict(weights["init"])
    sd.pop("conv_pre.bias")
    with pytest.raises(BackendUnavailableError):
        HiFTVocoder(sd, operand="fp32")
    with pytest.raises(ValueError):
        HiFTVocoder(weights["init"], operand="int8")


def test_full_size_batch_cfg2_properties(torch, weights, vocoders):
    """BASELINE.json configs[1] at full size (64 chunks x 500 frames, fp16 operands): size-independent properties
    for every chunk, oracle parity for one chunk drawn from the middle of the packed batch, and bit equality of
    another chunk with the same chunk run alone (tiling / packing independence)."""
    B, T = 64, 500
    L = 480 * T
    voc = vocoders("unit", "fp16")
    mels = [H.synth_mel(T, 1001, b) for b in range(B)]
    mel, Ts = voc.pack_mels(mels)
    f0 = torch.cat([H.synth_f0(T, 1001, b).reshape(-1) for b in range(B)]).cuda().contiguous()
    g = torch.Generator(device="cuda").manual_seed(2001)
    pv = (torch.rand(B, 9, generator=g, device="cuda") * 2 - 1) * math.pi
    pv[:, 0] = 0
    noise = torch.randn(B * 9 * L, generator=g, device="cuda")
    wav = voc.forward_packed(mel, Ts, f0=f0, phase_vec=pv.contiguous(), noise=noise).clone()
    torch.cuda.synchronize()
    assert wav.numel() == B * L and bool(torch.isfinite(wav).all())
    w = wav.view(B, L)
    assert float(w.abs().max()) <= 0.99 + 1e-7
    assert float(w[:, :480].abs().max()) == 0.0                      # trim_fade zeroes the first 20 ms of every chunk
    assert float(w[:, 960:].abs().amax(dim=1).min()) > 1e-3          # every chunk carries signal
    # determinism
    again = voc.forward_packed(mel, Ts, f0=f0, phase_vec=pv.contiguous(), noise=noise)
    assert torch.equal(wav, again)
    # one chunk against the oracle
    b = 37
    nz_b = noise[b * 9 * L:(b + 1) * 9 * L].view(9, L).cpu()
    ref = _oracle(torch, H.fold_weight_norm(weights["unit"]), mels[b], f0[b * T:(b + 1) * T].cpu(), pv[b].cpu(), nz_b)
    got = w[b].cpu()
    assert float((got - ref).abs().max()) <= MAX_ABS and H.snr_db(ref, got) >= MIN_SNR_DB, H.snr_db(ref, got)
    # one chunk alone == the same chunk inside the batch
    b = 5
    solo = voc.forward_packed(mel[b * T:(b + 1) * T].contiguous(), Ts[b:b + 1], f0=f0[b * T:(b + 1) * T].contiguous(),
                              phase_vec=pv[b:b + 1].contiguous(), noise=noise[b * 9 * L:(b + 1) * 9 * L].contiguous())
    assert torch.equal(solo, w[b])


def test_bf16_operand_mode_runs_at_its_documented_accuracy(torch, weights):
    """VT_OPERAND_BF16 is kept as a measured, NOT parity-green mode (an 8-bit mantissa cannot hold
    60 dB through 72 sequential convs).  The test pins what it does deliver so that the bf16 template instances stay
    exercised: finite output of the right length at >= 40 dB."""
    from vocalie_tts_b200.hift import HiFTVocoder
    Ts = [64, 9]
    mels, f0s, pvs, nzs = _inputs(torch, Ts, seed=21)
    voc = HiFTVocoder(weights["unit"], operand="bf16")
    wavs = voc.inference(mels, f0=f0s, phase_vec=pvs, noise=nzs)
    W = H.fold_weight_norm(weights["unit"])
    for b, T in enumerate(Ts):
        ref = _oracle(torch, W, mels[b], f0s[b], pvs[b], nzs[b])
        got = wavs[b].cpu()
        assert got.numel() == 480 * T and bool(torch.isfinite(got).all())
        assert H.snr_db(ref, got) >= 40.0, (b, H.snr_db(ref, got))


@pytest.mark.parametrize("env", [
    {"VT_CONVT": "0", "VT_PAIR_TR": "0", "VT_PAIR64": "0", "VT_CONV_2CTA": "0"},
    {"VT_PAIR64": "all"},
    {"VT_PAIR64": "all", "VT_P64_NA1": "1"},
    {"VT_PAIR_MC": "1"},
    {"VT_WSCALE": "1"},
    {"VT_WSCALE": "1", "VT_CONVT": "0", "VT_PAIR_TR": "0", "VT_PAIR64": "0"},
    {"VT_PAIR3": "0"},
    {"VT_PAIR3": "0", "VT_PAIR_TR": "0"},
    {"VT_PAIR3": "64"},
    {"VT_TC_DBG": "1024"},
    {"VT_PAIR3_ORDER": "0"},
], ids=["activation-major", "tap-paired-all", "tap-paired-single-a1", "weight-multicast", "row-scaled-weights",
        "row-scaled-activation-major", "separate-last-pairs", "separate-last-pairs-activation-major",
        "mean-fused-c64-only", "general-epilogue-paths", "mean-fused-ascending-kernel-sizes"])
def test_alternative_kernels_meet_the_parity_bar(env):
    """The kernel selections that are read from the environment once per process (VT_CONVT=0: activation-resident conv
    at C = 256, VT_PAIR_TR=0: untransposed pair kernel at C = 128, VT_PAIR64=0 / all: tap-paired pair kernel at C = 64
    for no / every pair, VT_P64_NA1=1: its single-A1 configuration, VT_PAIR_MC=1: weight ring multicast across a 2-CTA
    cluster, VT_WSCALE=1: power-of-two weight-row scales on EVERY layer - by default only layers with out-of-range rows
    get them, VT_PAIR3=0: the last pairs of a stage's three ResBlocks as three launches with a running mean in HBM
    instead of the mean-fused launch, VT_PAIR3=64: mean-fused at C = 64 only, VT_TC_DBG=1024: the transposed epilogues
    without their full-block fast paths, VT_PAIR3_ORDER=0: the subs of the C = 128 mean-fused launch in ascending kernel size) are alternative implementations of the same arithmetic; run them in a fresh process and hold them to the
    same waveform bar."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import torch\n"
        "from oracle import hift_oracle as H\n"
        "from vocalie_tts_b200.hift import HiFTVocoder\n"
        "sd = H.make_state_dict(0, 'unit'); W = H.fold_weight_norm(sd)\n"
        "voc = HiFTVocoder(sd, operand='fp16')\n"
        "Ts = [70, 11]\n"
        "mels = [H.synth_mel(T, 5, b) for b, T in enumerate(Ts)]\n"
        "f0s = [H.synth_f0(T, 5, b) for b, T in enumerate(Ts)]\n"
        "pn = [H.synth_noise(T, 5, b) for b, T in enumerate(Ts)]\n"
        "w = voc.inference(mels, f0=f0s, phase_vec=[p for p, _ in pn], noise=[n for _, n in pn])\n"
        "for b in range(len(Ts)):\n"
        "    ref = H.hift_inference(mels[b], W, f0=f0s[b], phase_vec=pn[b][0], noise=pn[b][1])\n"
        "    got = w[b].cpu()\n"
        "    assert got.numel() == ref.numel()\n"
        "    assert float((got - ref).abs().max()) <= 1e-3 and H.snr_db(ref, got) >= 60.0, H.snr_db(ref, got)\n"
        "print('fallback-ok')\n"
    ) % str(root)
    env = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fallback-ok" in r.stdout, r.stdout + r.stderr
