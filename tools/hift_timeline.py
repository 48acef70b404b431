"""Per-section device timeline of one HiFT forward at bench size (GPU box).  Honors VT_TC_DBG
(timing ablation of the tensor-core kernel; results are wrong with it) and prints one JSON line."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vocalie_tts_b200.hift import HiFTVocoder, random_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=64)
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--operand", default="fp16")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    voc = HiFTVocoder(random_state_dict(0), operand=a.operand)
    g = torch.Generator().manual_seed(1001)
    mels = [(torch.randn(80, a.frames, generator=g) * 2.0 - 5.0).clamp(-11.5129, 2.0) for _ in range(a.chunks)]
    mel, T = voc.pack_mels(mels)
    voc.set_profiling(True)
    best = None
    for _ in range(a.reps):
        voc.forward_packed(mel, T, seed=1)
        torch.cuda.synchronize()
        tl = voc.read_timeline()
        if best is None or sum(tl.values()) < sum(best.values()):
            best = tl
    best["total"] = sum(best.values())
    print(json.dumps({"dbg": int(os.environ.get("VT_TC_DBG", "0")), "chunks": a.chunks, "frames": a.frames,
                      "ms": {k: round(v, 3) for k, v in best.items()}}))


if __name__ == "__main__":
    main()
