"""GPU parity of the post-processing path (through the C ABI) against the pinned oracle and the
golden vectors produced by the reference's own functions.  Bit-exact: trim indices, lengths,
float32 sample bit patterns and PCM_16 codes."""
import json
import wave

import numpy as np
import pytest

from oracle import post_oracle as po

pytestmark = pytest.mark.gpu
SR = 24000


@pytest.fixture(scope="module")
def vt():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import vocalie_tts_b200 as v
    v.load_library()
    from vocalie_tts_b200 import post
    return post


def _speechlike(rng, n, lead, tail, amp=0.3, floor=0.0015):
    x = (rng.standard_normal(n) * amp).astype(np.float32)
    np.clip(x, -0.99, 0.99, out=x)
    x[:lead] = rng.uniform(-floor, floor, lead).astype(np.float32)
    if tail:
        x[n - tail:] = rng.uniform(-floor, floor, tail).astype(np.float32)
    return x


def test_reference_kats(vt):
    # reference tests/test_audio_edges.py:6-27
    assert vt._snap_zero_crossing(np.array([0.5, -0.2, 0.0, 0.3], np.float32), 3, radius_samples=3) == 3
    assert vt._fade_out(np.ones(10, np.float32), 5)[-1] == 0.0
    assert vt._fade_in(np.ones(10, np.float32), 5)[0] == 0.0
    a = np.array([0.0, 0.0, 0.01, 0.02, 0.0, 0.0], np.float32)
    assert vt._find_active_range(a, threshold=0.005, min_silence_frames=0) == (2, 4)


def test_ranges_and_snaps_match_golden(vt, golden):
    for name in golden.index["cases"]:
        x = golden[f"in_{name}"]
        far = vt._find_active_range(x, threshold=0.002, min_silence_frames=480)
        far0 = vt._find_active_range(x, threshold=0.005, min_silence_frames=0)
        s = vt._snap_zero_crossing(x, far[0], radius_samples=240)
        e = vt._snap_zero_crossing(x, max(far[1] - 1, s), radius_samples=240) + 1
        assert tuple(golden[f"far_{name}"]) == far + far0 + (s, e), name
        snaps = [vt._snap_zero_crossing(x, int(i), radius_samples=int(r))
                 for i, r in [(0, 240), (x.size // 2, 240), (x.size - 1, 240), (x.size + 5, 3), (3, 3), (x.size // 3, 17)]]
        assert list(golden[f"snaps_{name}"]) == snaps, name


def test_ramps_bit_exact(vt, golden):
    for f in (1, 2, 3, 5, 240, 100):
        a = np.ones(f + 7, np.float32)
        assert np.array_equal(golden[f"fin_{f}"].view(np.uint32), vt._fade_in(a.copy(), f).view(np.uint32)), f
        assert np.array_equal(golden[f"fout_{f}"].view(np.uint32), vt._fade_out(a.copy(), f).view(np.uint32)), f


def _write_wav(path, x):
    q = po.pcm16_encode(x)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(SR)
        w.writeframes(q.astype("<i2").tobytes())


def _read_wav(path):
    with wave.open(str(path), "rb") as w:
        return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int16)


def test_minimal_post_process_files_match_golden(vt, golden, tmp_path):
    for name in golden.index["file_cases"]:
        raw = tmp_path / f"{name}_raw.wav"
        out = tmp_path / f"{name}_out.wav"
        _write_wav(raw, golden[f"in_{name}"])
        meta = vt.minimal_post_process(raw, out)
        ref = golden.meta(f"mpp_meta_{name}")
        assert meta["trim"] == ref["trim"], name
        assert meta["peak_before"] == ref["peak_before"], name
        assert meta["normalize_scale"] == ref["normalize_scale"], name
        assert set(meta) == set(ref), name
        assert np.array_equal(_read_wav(out), golden[f"mpp_pcm_{name}"]), name


@pytest.mark.parametrize("tag,kw", [
    ("tn", dict(trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0)),
    ("t", dict(trim_enabled=True, normalize_enabled=False, target_dbfs=-3.0)),
    ("n", dict(trim_enabled=False, normalize_enabled=True, target_dbfs=-6.0)),
])
def test_apply_minimal_edit_files_match_golden(vt, golden, tmp_path, tag, kw):
    for name in golden.index["file_cases"]:
        if f"ame_{tag}_pcm_{name}" not in golden.z:
            continue
        raw = tmp_path / f"{name}_raw.wav"
        out = tmp_path / f"{name}_{tag}.wav"
        _write_wav(raw, golden[f"in_{name}"])
        res = vt.apply_minimal_edit(raw, out, **kw)
        ref = golden.meta(f"ame_{tag}_meta_{name}")
        for k in ("trimmed", "normalized", "peak_before", "peak_after", "gain", "target_dbfs"):
            assert res[k] == ref[k], (name, tag, k, res[k], ref[k])
        assert np.array_equal(_read_wav(out), golden[f"ame_{tag}_pcm_{name}"]), (name, tag)


def test_gap_stitch_matches_golden(vt, golden):
    for sname in golden.index["stitch"]:
        lens = golden[f"st_n_{sname}"]
        flat = golden[f"st_in_{sname}"]
        off = np.concatenate([[0], np.cumsum(lens)])
        chunks = [flat[off[i]:off[i + 1]] for i in range(len(lens))]
        for gap in (0, 250, 10, 2000):
            y = vt._apply_inter_chunk_gap(chunks, sr=SR, gap_ms=gap)
            ref = golden[f"st_out_{sname}_{gap}"]
            assert y.size == ref.size == po.stitched_length(lens, sr=SR, gap_ms=gap), (sname, gap)
            assert np.array_equal(y.view(np.uint32), ref.view(np.uint32)), (sname, gap)


def test_same_file_rejected(vt, tmp_path):
    p = tmp_path / "a.wav"
    _write_wav(p, np.zeros(100, np.float32))
    with pytest.raises(ValueError):
        vt.minimal_post_process(p, p)
    with pytest.raises(ValueError):
        vt.apply_minimal_edit(p, p, trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0)


def test_empty_and_tiny(vt, tmp_path):
    assert vt._find_active_range(np.zeros(0, np.float32), threshold=0.002, min_silence_frames=480) == (0, 0)
    assert vt._snap_zero_crossing(np.zeros(0, np.float32), 7, radius_samples=3) == 7
    assert vt._apply_inter_chunk_gap([], sr=SR, gap_ms=250).size == 0
    raw = tmp_path / "e.wav"
    out = tmp_path / "e_out.wav"
    _write_wav(raw, np.zeros(0, np.float32))
    meta = vt.minimal_post_process(raw, out)
    assert meta["trim"] == {"start_sample": 0, "end_sample": 0}
    assert _read_wav(out).size == 0
    # empty chunks inside a stitch
    rng = np.random.default_rng(3)
    chunks = [_speechlike(rng, 900, 5, 5), np.zeros(0, np.float32), _speechlike(rng, 100, 0, 0)]
    y = vt._apply_inter_chunk_gap(chunks, sr=SR, gap_ms=250)
    ref = po.apply_inter_chunk_gap(chunks, sr=SR, gap_ms=250)
    assert np.array_equal(y.view(np.uint32), ref.view(np.uint32))


def test_batched_segments_vs_oracle(vt):
    """Ragged batch of segments through one analyse+write (the per-chunk granularity of the
    north star): every segment must equal the oracle's minimal_post_process_array."""
    import torch
    rng = np.random.default_rng(11)
    lens = [1, 5, 479, 480, 481, 960, 2047, 2048, 2049, 24000, 120000, 7, 100001, 33333]
    chunks = [_speechlike(rng, n, min(n // 3, int(rng.integers(0, 1500))), min(n // 3, int(rng.integers(0, 1500))))
              for n in lens]
    seg_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = torch.from_numpy(np.concatenate(chunks)).cuda()
    prm = vt.make_params(trim=1, min_silence_frames=480, snap_radius=240, fade_in_frames=240, fade_out_frames=240,
                         normalize=1, target_peak=float(10 ** (-1.0 / 20.0)), concat=1)
    r = vt.post_process_device(flat, seg_off, prm)
    out = r.out[:r.total].cpu().numpy()
    pos = 0
    for i, c in enumerate(chunks):
        y, meta = po.minimal_post_process_array(c, SR)
        assert (int(r.results[i, 0]), int(r.results[i, 1])) == (meta["trim"]["start_sample"], meta["trim"]["end_sample"]), i
        assert int(r.results[i, 5]) == y.size, i
        assert float(r.results[i, 2]) == meta["peak_before"], i
        assert float(r.results[i, 3]) == meta["normalize_scale"], i
        assert np.array_equal(out[pos:pos + y.size].view(np.uint32), y.view(np.uint32)), i
        pos += y.size
    assert pos == r.total


def test_rms_helper_matches_reference_definition(vt):
    """sqrt(mean(float64(x)^2)) - the reference's clip-validation helper (tts_backends/cosyvoice_backend.py:103,
    tests/test_qwen3_runner.py:58).  float64 accumulation in another order than numpy's pairwise sum: equal to
    1e-13 relative (stated tolerance; the reduction itself is deterministic, checked by running it twice)."""
    rng = np.random.default_rng(23)
    lens = [0, 1, 3, 4, 5, 1023, 24000, 100001, 7, 480000]
    chunks = [(rng.standard_normal(n) * 0.3).astype(np.float32) for n in lens]
    seg_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = np.concatenate(chunks)
    got = vt.rms_segments(flat, seg_off)
    again = vt.rms_segments(flat, seg_off)
    assert np.array_equal(got.view(np.uint64), again.view(np.uint64))
    for i, c in enumerate(chunks):
        want = po.rms(c)
        assert abs(got[i] - want) <= 1e-13 * max(want, 1e-30), (i, got[i], want)
    assert vt.rms(np.zeros(0, np.float32)) == 0.0
    assert vt.rms(np.full(1000, 0.5, np.float32)) == 0.5        # reference test value: a constant clip


def test_full_size_properties(vt):
    """cfg3-sized stitch (512 chunks x 10 s, gap 250 ms): size-independent properties -
    exact length, gaps are zero, interior samples untouched, idempotent trim."""
    import torch
    n_chunks, n = 512, 240000
    g = torch.Generator(device="cuda").manual_seed(7)
    flat = (torch.randn(n_chunks * n, generator=g, device="cuda") * 0.3).clamp_(-0.99, 0.99)
    seg_off = np.arange(n_chunks + 1, dtype=np.int64) * n
    prm = vt.stitch_params(n_chunks, sr=SR, gap_ms=250)
    r = vt.post_process_device(flat, seg_off, prm)
    assert r.total == po.stitched_length([n] * n_chunks, sr=SR, gap_ms=250) == 125_946_000
    out = r.out[:r.total]
    view = out[: (n + 6000) * (n_chunks - 1)].view(n_chunks - 1, n + 6000)
    assert float(view[:, n:].abs().max()) == 0.0                      # gaps are silent
    assert torch.equal(view[:, 240:n - 240], flat[: n * (n_chunks - 1)].view(n_chunks - 1, n)[:, 240:n - 240])
    assert float(view[1:, 0].abs().max()) == 0.0                      # fade-in starts at 0
    assert float(view[:, n - 1].abs().max()) == 0.0                   # fade-out ends at 0
    # spot-check three chunks bit-exactly against the oracle
    for i in (0, 255, 511):
        c = flat[i * n:(i + 1) * n].cpu().numpy()
        ref = c.copy()
        if i < n_chunks - 1:
            po.fade_out(ref, 240)
        if i > 0:
            po.fade_in(ref, 240)
        got = out[i * (n + 6000): i * (n + 6000) + n].cpu().numpy()
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), i


def test_library_errors_are_backend_errors(vt):
    import torch
    from vocalie_tts_b200 import BackendUnavailableError, _lib
    lib = _lib.load_library()
    x = torch.zeros(16, device="cuda")
    with pytest.raises(BackendUnavailableError):
        _lib.check(lib.vt_post_analyze(x.data_ptr(), 0, 1, 16, 16, None, 0, 0, 0, 0), "vt_post_analyze")
