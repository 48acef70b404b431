"""Prints the F0-predictor error of the tensor-core path against the fp32 oracle (debugging sessions; honours VT_LIB_PATH)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import hift_oracle as H
from vocalie_tts_b200.hift import HiFTVocoder
for kind in ("unit", "init"):
    sd = H.make_state_dict(0, kind); W = H.fold_weight_norm(sd)
    voc = HiFTVocoder(sd, operand="fp16")
    Ts = [300, 17, 256]
    mels = [H.synth_mel(T, 11, b) for b, T in enumerate(Ts)]
    pn = [H.synth_noise(T, 11, b) for b, T in enumerate(Ts)]
    voc.inference(mels, phase_vec=[p for p, _ in pn], noise=[n for _, n in pn])
    for b, T in enumerate(Ts):
        want = H.f0_predictor(mels[b].unsqueeze(0), W)[0].double()
        want64 = H.f0_predictor(mels[b].unsqueeze(0).double(), {k: v.double() for k, v in W.items()})[0]
        got = voc.read_tap("f0", b, 1).cpu().reshape(-1).double()
        sc = float(want.abs().max())
        print(kind, T, "vs fp32 oracle %.3e  vs fp64 oracle %.3e  (fp32 oracle vs fp64 %.3e)" % (
            float((got - want).abs().max()) / sc, float((got - want64).abs().max()) / sc, float((want - want64).abs().max()) / sc))
