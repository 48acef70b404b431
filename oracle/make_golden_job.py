"""Generate ``tests/golden/job_golden.npz``: whole-JOB post-processing as the reference runs it.

TEST INFRASTRUCTURE.  Run in the build container only (``/root/reference`` does not exist on the GPU
box):

    python oracle/make_golden_job.py

The reference does not post-process chunk by chunk.  ``run_tts_pipeline`` stitches the RAW chunks with
``_apply_inter_chunk_gap`` and writes the file with the soundfile default subtype (PCM_16)
(backend/shared/tts_pipeline.py:395-409); ``run_tts_job`` then runs ``apply_minimal_edit`` ONCE on that
file - whole-file trim, ONE peak, clip, PCM_16 (backend/services/tts_service.py:195-207,
backend/shared/audio_edit.py:16-79); the legacy surface runs ``minimal_post_process`` on the same file
(backend/shared/tts_pipeline.py:212-274).  This script drives exactly those reference functions
(imported unchanged; ``librosa`` stubbed, ``soundfile`` -> ``oracle/sf_stub.py``) on seeded chunk sets
and stores the inputs, the stitched raw file and every edited file as PCM_16 codes plus the returned
dicts.
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.make_golden import import_reference, speechlike, OUT  # noqa: E402


def main():
    tp, ae = import_reference()
    sf = sys.modules["soundfile"]
    rng = np.random.default_rng(20261018)
    sr = 24000
    g = {}
    jobs = {
        # loudness differs between chunks: a per-chunk normalise would change their balance
        "balance": [speechlike(rng, n, l, t, amp=a) for n, l, t, a in
                    ((9000, 1200, 700, 0.05), (14000, 300, 900, 0.3), (6000, 800, 2000, 0.12))],
        # pauses inside the job must survive (only the file's own head/tail are trimmed)
        "pauses": [speechlike(rng, n, l, t) for n, l, t in ((8000, 3000, 2500), (5000, 2000, 2200), (7000, 1500, 3000), (4000, 900, 1200))],
        "single": [speechlike(rng, 12000, 1000, 1400)],
        "short": [speechlike(rng, n, 0, 0) for n in (100, 479, 240, 1, 241, 3000)],
        "silent_edges": [np.zeros(3000, np.float32), speechlike(rng, 9000, 600, 600), np.zeros(2000, np.float32)],
        "all_silent": [np.zeros(2500, np.float32), np.zeros(1800, np.float32)],
        "many": [speechlike(rng, int(rng.integers(2400, 24000)), int(rng.integers(0, 900)), int(rng.integers(0, 900)),
                            amp=float(rng.uniform(0.05, 0.35))) for _ in range(13)],
    }
    edits = {"tn": dict(trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0),
             "t": dict(trim_enabled=True, normalize_enabled=False, target_dbfs=-3.0),
             "n": dict(trim_enabled=False, normalize_enabled=True, target_dbfs=-6.0)}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for name, chunks in jobs.items():
            g[f"n_{name}"] = np.array([c.size for c in chunks], np.int64)
            g[f"in_{name}"] = np.concatenate(chunks)
            for gap in (250, 0):
                # tts_pipeline.py:398-409
                if len(chunks) > 1 and gap > 0:
                    final = tp._apply_inter_chunk_gap([c.copy() for c in chunks], sr=sr, gap_ms=gap)
                else:
                    final = np.concatenate(chunks)
                raw = td / f"{name}_{gap}_raw.wav"
                sf.write(str(raw), final, sr)
                q, _ = sf.read(str(raw), dtype="int16")
                g[f"raw_{name}_{gap}"] = q
                for tag, kw in edits.items():
                    out = td / f"{name}_{gap}_{tag}.wav"
                    res = ae.apply_minimal_edit(raw, out, silence_threshold=0.002, silence_min_ms=20, **kw)
                    q2, _ = sf.read(str(out), dtype="int16")
                    g[f"ame_{tag}_{name}_{gap}"] = q2
                    g[f"ame_{tag}_meta_{name}_{gap}"] = np.frombuffer(json.dumps(res).encode(), np.uint8)
                out = td / f"{name}_{gap}_mpp.wav"
                meta = tp.minimal_post_process(raw, out)
                q3, _ = sf.read(str(out), dtype="int16")
                g[f"mpp_{name}_{gap}"] = q3
                g[f"mpp_meta_{name}_{gap}"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    g["cases"] = np.frombuffer(json.dumps({"jobs": list(jobs), "gaps": [250, 0], "edits": list(edits)}).encode(), np.uint8)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "job_golden.npz", **g)
    print(f"wrote {OUT / 'job_golden.npz'} with {len(g)} arrays")


if __name__ == "__main__":
    main()
