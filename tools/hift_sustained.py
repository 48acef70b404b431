"""Sustained (power-capped) forward time: 30 back-to-back forwards at bench size after 6 warm-ups, with the SM clock
and board power sampled while the queue is still busy.  B200 boards reach the 1 kW cap on this workload, so the
first forwards of a process run ~12 % faster than the steady state bench.py reports.  Honors VT_TC_DBG."""
import sys, os, subprocess
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vocalie_tts_b200.hift import HiFTVocoder, random_state_dict
voc = HiFTVocoder(random_state_dict(0), operand="fp16")
g = torch.Generator().manual_seed(1001)
mels = [(torch.randn(80, 500, generator=g) * 2.0 - 5.0).clamp(-11.5129, 2.0) for _ in range(64)]
mel, T = voc.pack_mels(mels)
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
for i in range(6):
    voc.forward_packed(mel, T, seed=i)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30):
    voc.forward_packed(mel, T, seed=i)
e1.record()
s = smi()
torch.cuda.synchronize()
print("dbg", os.environ.get("VT_TC_DBG", "0"), "sustained ms/forward %.2f" % (e0.elapsed_time(e1) / 30), "| clocks,power:", s)
