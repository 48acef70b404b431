"""Generate ``tests/golden/post_*.npz`` by running the REFERENCE's own functions.

TEST INFRASTRUCTURE.  Run in the build container only (``/root/reference`` does not
exist on the GPU box):

    python oracle/make_golden.py            # rewrites tests/golden/post_golden.npz

The reference modules are imported unchanged from ``/root/reference`` with two
``sys.modules`` stubs (``librosa`` - unused on this path; ``soundfile`` ->
``oracle/sf_stub.py``).  Inputs are seeded; both inputs and the reference's outputs are
stored so the tests never need the reference at run time.
"""
from __future__ import annotations

import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"


def import_reference():
    """Import the reference's post-processing modules (stubs for absent deps)."""
    if not REF.exists():
        raise RuntimeError("/root/reference is not available (GPU box?) - goldens are generated in the build container")
    sys.path.insert(0, str(ROOT))
    from oracle import sf_stub
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    sys.modules["soundfile"] = sf_stub.make_module()
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import backend.shared.tts_pipeline as tp  # noqa: E402
    import backend.shared.audio_edit as ae    # noqa: E402
    return tp, ae


def speechlike(rng, n, lead, tail, amp=0.3, floor=0.0015):
    """Near-silent head/tail (|x| < 0.002) around a clipped-gaussian body."""
    x = (rng.standard_normal(n) * amp).astype(np.float32)
    np.clip(x, -0.99, 0.99, out=x)
    x[:lead] = (rng.uniform(-floor, floor, lead)).astype(np.float32)
    if tail:
        x[n - tail:] = (rng.uniform(-floor, floor, tail)).astype(np.float32)
    return x


def main():
    tp, ae = import_reference()
    rng = np.random.default_rng(20251018)
    sr = 24000
    g = {}
    cases = []

    # ---- find_active_range / snap / minimal_post_process semantics on arrays
    inputs = []
    inputs.append(("kat_edges", np.array([0.0, 0.0, 0.01, 0.02, 0.0, 0.0], np.float32)))
    inputs.append(("kat_snap", np.array([0.5, -0.2, 0.0, 0.3], np.float32)))
    inputs.append(("silent", np.zeros(4800, np.float32)))
    inputs.append(("one_sample", np.array([0.5], np.float32)))
    inputs.append(("thr_exact", np.array([0.0] * 600 + [np.float32(0.002)] * 3 + [0.0021] + [0.0] * 700, np.float32)))
    inputs.append(("short_active", np.concatenate([np.zeros(700, np.float32), np.float32([0.2, -0.3, 0.1]), np.zeros(900, np.float32)])))
    inputs.append(("square_0p2s", np.concatenate([np.zeros(1200, np.float32), np.full(2400, 0.1, np.float32), np.zeros(1200, np.float32)])))
    for i in range(12):
        n = int(rng.integers(2000, 24000))
        lead = int(rng.integers(0, 1500))
        tail = int(rng.integers(0, 1500))
        inputs.append((f"speech{i}", speechlike(rng, n, lead, tail)))
    # a 5 s chunk like config 1
    inputs.append(("cfg1_like", speechlike(rng, 120000, 2400, 3100)))
    # sparse signal: many exact zeros around the boundaries (zero-valued neighbours count as crossings)
    z = speechlike(rng, 30000, 900, 900)
    z[rng.integers(0, 30000, 6000)] = 0.0
    inputs.append(("zeros_sprinkled", z))
    # all-positive signal: no sign change -> snap falls back to idx
    inputs.append(("all_positive", np.abs(speechlike(rng, 20000, 800, 800, floor=0.0)) + np.float32(1e-4)))

    for name, x in inputs:
        g[f"in_{name}"] = x
        far = tp._find_active_range(x, threshold=0.002, min_silence_frames=480)
        far0 = tp._find_active_range(x, threshold=0.005, min_silence_frames=0)
        s = tp._snap_zero_crossing(x, far[0], radius_samples=240)
        e = tp._snap_zero_crossing(x, max(far[1] - 1, s), radius_samples=240) + 1
        snaps = [tp._snap_zero_crossing(x, int(i), radius_samples=int(r))
                 for i, r in [(0, 240), (x.size // 2, 240), (x.size - 1, 240), (x.size + 5, 3), (3, 3), (x.size // 3, 17)]]
        g[f"far_{name}"] = np.array(far + far0 + (s, e), np.int64)
        g[f"snaps_{name}"] = np.array(snaps, np.int64)
        cases.append(name)

    # ---- file-level functions (through the soundfile stub: PCM_16 in, PCM_16 out)
    sf = sys.modules["soundfile"]
    file_cases = ["square_0p2s", "speech0", "speech3", "cfg1_like", "zeros_sprinkled", "silent", "all_positive"]
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for name in file_cases:
            x = g[f"in_{name}"]
            raw = td / f"{name}_raw.wav"
            sf.write(str(raw), x, sr)
            out = td / f"{name}_mpp.wav"
            meta = tp.minimal_post_process(raw, out)
            q, _ = sf.read(str(out), dtype="int16")
            g[f"mpp_pcm_{name}"] = q
            g[f"mpp_meta_{name}"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
            for tag, kw in [("tn", dict(trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0)),
                            ("t", dict(trim_enabled=True, normalize_enabled=False, target_dbfs=-3.0)),
                            ("n", dict(trim_enabled=False, normalize_enabled=True, target_dbfs=-6.0))]:
                if name == "cfg1_like" and tag != "tn":
                    continue
                out2 = td / f"{name}_ame_{tag}.wav"
                res = ae.apply_minimal_edit(raw, out2, **kw)
                q2, _ = sf.read(str(out2), dtype="int16")
                g[f"ame_{tag}_pcm_{name}"] = q2
                g[f"ame_{tag}_meta_{name}"] = np.frombuffer(json.dumps(res).encode(), np.uint8)

    # ---- gap stitching
    stitch_sets = {
        "three": [speechlike(rng, int(n), 100, 100) for n in (5000, 7000, 3000)],
        "short": [speechlike(rng, int(n), 0, 0) for n in (100, 479, 240, 1, 241, 3000)],  # shorter than 2*fade
        "single": [speechlike(rng, 4000, 10, 10)],
        "many": [speechlike(rng, int(rng.integers(2400, 24000)), 50, 50) for _ in range(8)],
    }
    for sname, chunks in stitch_sets.items():
        g[f"st_n_{sname}"] = np.array([c.size for c in chunks], np.int64)
        g[f"st_in_{sname}"] = np.concatenate(chunks)
        for gap in (0, 250, 10, 2000):
            y = tp._apply_inter_chunk_gap([c.copy() for c in chunks], sr=sr, gap_ms=gap)
            g[f"st_out_{sname}_{gap}"] = y

    # ---- ramps (bit patterns of np.linspace as the reference builds them)
    for f in (1, 2, 3, 5, 240, 100):
        a = np.ones(f + 7, np.float32)
        g[f"fin_{f}"] = tp._fade_in(a.copy(), f)
        g[f"fout_{f}"] = tp._fade_out(a.copy(), f)

    g["cases"] = np.frombuffer(json.dumps({"cases": cases, "file_cases": file_cases,
                                           "stitch": list(stitch_sets)}).encode(), np.uint8)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "post_golden.npz", **g)
    print(f"wrote {OUT / 'post_golden.npz'} with {len(g)} arrays")


if __name__ == "__main__":
    main()
