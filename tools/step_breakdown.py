"""Where the bench step's time goes beyond the vocoder forward: device time of (a) the forward alone, (b) the job post
alone (stitch -> PCM_16 -> decode -> whole-file edit) on a resident waveform, (c) run_device (both), each as the mean of
back-to-back calls between two events, with and without the library's profiling events.  GPU box only."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vocalie_tts_b200.hift import HiFTVocoder, random_state_dict  # noqa: E402
from vocalie_tts_b200.pipeline import VocoderPipeline  # noqa: E402

voc = HiFTVocoder(random_state_dict(0), operand="fp16")
pipe = VocoderPipeline(voc, chunk_gap_ms=250, out_pcm16=True)
g = torch.Generator().manual_seed(1001)
mels = [(torch.randn(80, 500, generator=g) * 2.0 - 5.0).clamp(-11.5129, 2.0) for _ in range(64)]
mel, T = voc.pack_mels(mels)
T = np.ascontiguousarray(T, dtype=np.int32)
n = int(T.astype(np.int64).sum()) * pipe.spf
seg_off = np.concatenate([[0], np.cumsum(T.astype(np.int64) * pipe.spf)])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=voc.device)


def timed(fn, reps, each_flush=False):
    torch.cuda.synchronize()
    tot = 0.0
    if each_flush:
        for i in range(reps):
            flush.fill_(i & 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(i); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for prof in (False, True):
    voc.set_profiling(prof)
    for i in range(8):
        pipe.run_device(mel, T, seed=i)
    wav = pipe._wav
    fwd = timed(lambda i: voc.forward_bucketed(mel, T, seed=i, out=pipe._buf("_wav", n + 4, torch.float32)), 20)
    post = timed(lambda i: pipe.post_device(wav, seg_off), 20)
    both = timed(lambda i: pipe.run_device(mel, T, seed=i), 20)
    step = timed(lambda i: pipe.run_device(mel, T, seed=i), 10, each_flush=True)
    print(f"profiling={prof}: forward {fwd:.3f} ms, post {post:.3f} ms, run_device {both:.3f} ms back to back; "
          f"{step:.3f} ms as bench.py times it (L2 flush + sync per step)")
