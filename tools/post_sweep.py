"""Post-processing-only sweep (BASELINE.json configs[4]): trim + snap + fades + peak normalise +
250 ms gap concat over N float32 samples split into 10 s segments.  Prints one JSON line with the
achieved HBM GB/s (12 B per input sample)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vocalie_tts_b200 import post  # noqa: E402


def make_input(n_total, seg=240000, seed=1005):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(n_total, generator=g, device="cuda") * 0.3).clamp_(-0.99, 0.99)
    n_seg = (n_total + seg - 1) // seg
    seg_off = np.minimum(np.arange(n_seg + 1, dtype=np.int64) * seg, n_total)
    cpu = torch.Generator().manual_seed(seed)
    lead = torch.randint(1200, 7200, (n_seg,), generator=cpu).tolist()
    tail = torch.randint(1200, 7200, (n_seg,), generator=cpu).tolist()
    for i in range(n_seg):
        a, b = int(seg_off[i]), int(seg_off[i + 1])
        la, ta = min(lead[i], (b - a) // 3), min(tail[i], (b - a) // 3)
        x[a:a + la] *= 0.004
        x[b - ta:b] *= 0.004
    return x, seg_off


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=268435456)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pcm16", type=int, default=0)
    a = ap.parse_args()
    x, seg_off = make_input(a.samples)
    n_seg = len(seg_off) - 1
    prm = post.make_params(trim=1, min_silence_frames=480, snap_radius=240, fade_in_frames=240, fade_out_frames=240,
                           normalize=1, target_peak=float(10 ** (-1 / 20)), stitch=0, gap_frames=0, concat=1,
                           out_pcm16=a.pcm16)
    cap = a.samples + n_seg * 6000
    out = torch.empty(cap, dtype=torch.int16 if a.pcm16 else torch.float32, device="cuda")
    for _ in range(a.warmup):
        post.post_process_device(x, seg_off, prm, out=out, read_back=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        post.post_process_device(x, seg_off, prm, out=out, read_back=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    bps = 10 if a.pcm16 else 12
    gbs = bps * a.samples / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": "post sweep", "samples": a.samples, "segments": n_seg, "ms": ms,
                      "bytes_per_sample": bps, "achieved_gbs": gbs, "peak_gbs": 6544.3, "frac": gbs / 6544.3,
                      "audio_s_per_s": a.samples / 24000 / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
