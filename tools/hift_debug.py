"""Layer-by-layer error table of the CUDA HiFT path against the oracle (debugging aid, GPU box)."""
import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import hift_oracle as H  # noqa: E402
from vocalie_tts_b200.hift import HiFTVocoder  # noqa: E402

TAP_CH = {"s": 1, "s_stft": 18, "conv_pre": 512, "ups0": 256, "x0": 256, "stage0": 256, "ups1": 128, "x1": 128,
          "stage1": 128, "ups2": 64, "x2": 64, "stage2": 64, "conv_post": 18}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="unit")
    ap.add_argument("--operand", default="fp32")
    ap.add_argument("--T", type=int, nargs="+", default=[37, 50, 8])
    a = ap.parse_args()
    sd = H.make_state_dict(0, a.kind)
    W = H.fold_weight_norm(sd)
    voc = HiFTVocoder(sd, operand=a.operand)
    mels = [H.synth_mel(T, 3, b) for b, T in enumerate(a.T)]
    f0s = [H.synth_f0(T, 3, b) for b, T in enumerate(a.T)]
    pn = [H.synth_noise(T, 3, b) for b, T in enumerate(a.T)]
    t0 = time.time()
    wavs = voc.inference(mels, f0=f0s, phase_vec=[p for p, _ in pn], noise=[n for _, n in pn])
    torch.cuda.synchronize()
    print(f"forward {time.time() - t0:.3f}s launches={voc.last_launches}")
    for b, T in enumerate(a.T):
        taps = {}
        ref = H.hift_inference(mels[b], W, f0=f0s[b], phase_vec=pn[b][0], noise=pn[b][1], taps=taps)
        for name, ch in TAP_CH.items():
            got = voc.read_tap(name, b, ch).cpu()
            want = taps[name][0].t().contiguous()
            if got.shape != want.shape:
                print(f"  seq{b} {name:10s} SHAPE {tuple(got.shape)} vs {tuple(want.shape)}")
                continue
            err = (got - want).abs()
            print(f"  seq{b} {name:10s} max|ref|={float(want.abs().max()):.4g} maxerr={float(err.max()):.3g} "
                  f"rel={float(err.max()) / (float(want.abs().max()) + 1e-12):.3g} argmax={int(err.argmax()) // ch},{int(err.argmax()) % ch}"
                  f" nan={int(torch.isnan(got).sum())}")
        got = wavs[b].cpu()
        print(f"  seq{b} WAV n={got.numel()} maxerr={float((got - ref).abs().max()):.3g} snr={H.snr_db(ref, got):.1f} dB")


if __name__ == "__main__":
    main()
