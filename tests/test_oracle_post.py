"""The numpy post-processing oracle vs the REFERENCE's own outputs (golden vectors)
and the reference's four known-answer tests (reference tests/test_audio_edges.py:6-27)."""
import numpy as np
import pytest

from oracle import post_oracle as po

SR = 24000


def test_reference_kat_snap():
    a = np.array([0.5, -0.2, 0.0, 0.3], np.float32)
    assert po.snap_zero_crossing(a, 3, radius_samples=3) == 3


def test_reference_kat_fades():
    assert po.fade_out(np.ones(10, np.float32), 5)[-1] == 0.0
    assert po.fade_in(np.ones(10, np.float32), 5)[0] == 0.0


def test_reference_kat_active_range():
    a = np.array([0.0, 0.0, 0.01, 0.02, 0.0, 0.0], np.float32)
    assert po.find_active_range(a, threshold=0.005, min_silence_frames=0) == (2, 4)


def test_ranges_and_snaps_match_reference(golden):
    for name in golden.index["cases"]:
        x = golden[f"in_{name}"]
        far = po.find_active_range(x, threshold=0.002, min_silence_frames=480)
        far0 = po.find_active_range(x, threshold=0.005, min_silence_frames=0)
        s = po.snap_zero_crossing(x, far[0], radius_samples=240)
        e = po.snap_zero_crossing(x, max(far[1] - 1, s), radius_samples=240) + 1
        assert tuple(golden[f"far_{name}"]) == far + far0 + (s, e), name
        snaps = [po.snap_zero_crossing(x, int(i), radius_samples=int(r))
                 for i, r in [(0, 240), (x.size // 2, 240), (x.size - 1, 240), (x.size + 5, 3), (3, 3), (x.size // 3, 17)]]
        assert list(golden[f"snaps_{name}"]) == snaps, name


def test_ramps_bit_exact(golden):
    for f in (1, 2, 3, 5, 240, 100):
        a = np.ones(f + 7, np.float32)
        assert np.array_equal(golden[f"fin_{f}"].view(np.uint32), po.fade_in(a.copy(), f).view(np.uint32))
        assert np.array_equal(golden[f"fout_{f}"].view(np.uint32), po.fade_out(a.copy(), f).view(np.uint32))


def test_minimal_post_process_matches_reference_files(golden):
    for name in golden.index["file_cases"]:
        raw = po.pcm16_decode(po.pcm16_encode(golden[f"in_{name}"]))  # what sf.read gives back
        y, meta = po.minimal_post_process_array(raw, SR)
        ref_meta = golden.meta(f"mpp_meta_{name}")
        assert meta["trim"] == ref_meta["trim"], name
        assert meta["peak_before"] == ref_meta["peak_before"], name
        assert meta["normalize_scale"] == ref_meta["normalize_scale"], name
        assert np.array_equal(po.pcm16_encode(y), golden[f"mpp_pcm_{name}"]), name


@pytest.mark.parametrize("tag,kw", [
    ("tn", dict(trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0)),
    ("t", dict(trim_enabled=True, normalize_enabled=False, target_dbfs=-3.0)),
    ("n", dict(trim_enabled=False, normalize_enabled=True, target_dbfs=-6.0)),
])
def test_apply_minimal_edit_matches_reference_files(golden, tag, kw):
    for name in golden.index["file_cases"]:
        if f"ame_{tag}_pcm_{name}" not in golden.z:
            continue
        raw = po.pcm16_decode(po.pcm16_encode(golden[f"in_{name}"]))
        y, res = po.apply_minimal_edit_array(raw, SR, **kw)
        ref = golden.meta(f"ame_{tag}_meta_{name}")
        for k in ("trimmed", "normalized", "peak_before", "peak_after", "gain", "target_dbfs"):
            assert res[k] == ref[k], (name, k)
        assert np.array_equal(po.pcm16_encode(y), golden[f"ame_{tag}_pcm_{name}"]), name


def test_gap_stitch_matches_reference(golden):
    for sname in golden.index["stitch"]:
        lens = golden[f"st_n_{sname}"]
        flat = golden[f"st_in_{sname}"]
        off = np.concatenate([[0], np.cumsum(lens)])
        chunks = [flat[off[i]:off[i + 1]] for i in range(len(lens))]
        for gap in (0, 250, 10, 2000):
            y = po.apply_inter_chunk_gap(chunks, sr=SR, gap_ms=gap)
            ref = golden[f"st_out_{sname}_{gap}"]
            assert y.size == ref.size == po.stitched_length(lens, sr=SR, gap_ms=gap), (sname, gap)
            assert np.array_equal(y.view(np.uint32), ref.view(np.uint32)), (sname, gap)


def test_empty_inputs():
    assert po.find_active_range(np.zeros(0, np.float32), threshold=0.002, min_silence_frames=480) == (0, 0)
    assert po.snap_zero_crossing(np.zeros(0, np.float32), 7, radius_samples=3) == 7
    assert po.apply_inter_chunk_gap([], sr=SR, gap_ms=250).size == 0
    y, meta = po.minimal_post_process_array(np.zeros(0, np.float32), SR)
    assert y.size == 0 and meta["trim"] == {"start_sample": 0, "end_sample": 0}
