"""GPU tests of the JOB-level post path in the reference's order (VocoderPipeline granularity="job"):
stitch raw chunks -> PCM_16 raw file -> one whole-file edit, against files produced by the REFERENCE's own
functions (tests/golden/job_golden.npz, oracle/make_golden_job.py); the one-pass ``normalize=2`` kernel mode;
two job threads hammering the post functions concurrently (backend/config.py:11 MAX_CONCURRENT_JOBS = 2)."""
import json
import threading
from pathlib import Path

import numpy as np
import pytest

from oracle import post_oracle as po

pytestmark = pytest.mark.gpu
SR = 24000
GOLD = Path(__file__).resolve().parent / "golden" / "job_golden.npz"


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import vocalie_tts_b200 as v
    v.load_library()
    z = np.load(GOLD)
    idx = json.loads(bytes(z["cases"]).decode())
    return torch, z, idx


class _Voc:
    def __init__(self, torch):
        self.device = torch.device("cuda", torch.cuda.current_device())


def _job(z, name):
    lens = z[f"n_{name}"]
    return z[f"in_{name}"], np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)


@pytest.mark.parametrize("tag", ["tn", "t", "n"])
def test_job_pipeline_matches_the_reference_files(env, tag):
    torch, z, idx = env
    from vocalie_tts_b200.pipeline import VocoderPipeline
    kw = {"tn": dict(trim_silence=True, normalize=True, target_dbfs=-1.0), "t": dict(trim_silence=True, normalize=False, target_dbfs=-3.0),
          "n": dict(trim_silence=False, normalize=True, target_dbfs=-6.0)}[tag]
    for gap in idx["gaps"]:
        pipe = VocoderPipeline(_Voc(torch), chunk_gap_ms=gap, out_pcm16=True, **kw)
        for name in idx["jobs"]:
            flat, seg_off = _job(z, name)
            res = pipe.post_device(torch.from_numpy(flat).cuda(), seg_off, read_back=True)
            got = res.audio[: res.total_samples].cpu().numpy()
            raw = res.raw[: res.raw_samples].cpu().numpy()
            assert np.array_equal(raw, z[f"raw_{name}_{gap}"]), (name, gap, "raw file")
            want = z[f"ame_{tag}_{name}_{gap}"]
            assert got.size == want.size, (name, gap, tag, got.size, want.size)
            assert np.array_equal(got, want), (name, gap, tag)
            meta = json.loads(bytes(z[f"ame_{tag}_meta_{name}_{gap}"]).decode())
            e = res.edit
            assert e["trimmed"] == meta["trimmed"] and e["normalized"] == meta["normalized"], (name, gap, tag)
            assert e["peak_before"] == meta["peak_before"] and e["gain"] == meta["gain"], (name, gap, tag)


def test_job_pipeline_minimal_post_process_variant(env):
    torch, z, idx = env
    from vocalie_tts_b200.pipeline import VocoderPipeline
    for gap in idx["gaps"]:
        pipe = VocoderPipeline(_Voc(torch), chunk_gap_ms=gap, out_pcm16=True, edit="minimal_post_process")
        for name in idx["jobs"]:
            flat, seg_off = _job(z, name)
            res = pipe.post_device(torch.from_numpy(flat).cuda(), seg_off, read_back=True)
            got = res.audio[: res.total_samples].cpu().numpy()
            assert np.array_equal(got, z[f"mpp_{name}_{gap}"]), (name, gap)
            meta = json.loads(bytes(z[f"mpp_meta_{name}_{gap}"]).decode())
            assert (res.edit["start_sample"], res.edit["end_sample"]) == (meta["trim"]["start_sample"], meta["trim"]["end_sample"])
            assert res.edit["peak_before"] == meta["peak_before"]


def test_job_without_edit_is_the_stitched_file(env):
    torch, z, idx = env
    from vocalie_tts_b200.pipeline import VocoderPipeline
    pipe = VocoderPipeline(_Voc(torch), chunk_gap_ms=250, trim_silence=False, normalize=False)
    flat, seg_off = _job(z, "many")
    res = pipe.post_device(torch.from_numpy(flat).cuda(), seg_off, read_back=True)
    lens = np.diff(seg_off)
    chunks = [flat[seg_off[i]:seg_off[i + 1]] for i in range(len(lens))]
    want = po.apply_inter_chunk_gap(chunks, sr=SR, gap_ms=250)
    got = res.audio[: res.total_samples].cpu().numpy()
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(po.pcm16_encode(got), z["raw_many_250"])


def test_normalize2_one_peak_over_all_segments(env):
    """Kernel mode normalize=2 (one pass: stitch + ONE peak + gain + clip): equals _apply_inter_chunk_gap followed by
    apply_minimal_edit's normalise on the float file (no PCM_16 round trip in between), bit for bit."""
    torch, z, idx = env
    from vocalie_tts_b200 import post
    for name in ("balance", "many", "short", "all_silent"):
        flat, seg_off = _job(z, name)
        chunks = [flat[seg_off[i]:seg_off[i + 1]] for i in range(len(seg_off) - 1)]
        prm = post.stitch_params(len(chunks), sr=SR, gap_ms=250, normalize=2, clip=1, target_peak=float(10 ** (-1.0 / 20.0)))
        r = post.post_process_device(torch.from_numpy(flat).cuda(), seg_off, prm)
        got = r.out[: r.total].cpu().numpy()
        stitched = po.apply_inter_chunk_gap(chunks, sr=SR, gap_ms=250)
        want, res = po.apply_minimal_edit_array(stitched, SR, trim_enabled=False, normalize_enabled=True, target_dbfs=-1.0)
        assert got.size == want.size and np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32)), name
        assert float(r.results[0, 6]) == res["peak_before"], name          # the peak every segment's gain used
        # peak_override (the cross-rank hook) replaces the analysed peak
        pk = torch.tensor([0.5], dtype=torch.float32, device="cuda")
        r2 = post.post_process_device(torch.from_numpy(flat).cuda(), seg_off, prm, peak_override=pk)
        want2 = np.clip(stitched * (float(10 ** (-1.0 / 20.0)) / 0.5), -1.0, 1.0).astype(np.float32)
        assert np.array_equal(r2.out[: r2.total].cpu().numpy().view(np.uint32), want2.view(np.uint32)), name


def test_post_stats_matches_numpy(env):
    torch, z, idx = env
    from vocalie_tts_b200 import post
    flat, seg_off = _job(z, "silent_edges")
    fl, pk = post.stats_device(torch.from_numpy(flat).cuda(), seg_off)
    fl, pk = fl.cpu().numpy(), pk.cpu().numpy()
    for i in range(len(seg_off) - 1):
        seg = flat[seg_off[i]:seg_off[i + 1]]
        act = np.flatnonzero(np.abs(seg) > np.float32(0.002))
        assert tuple(fl[i]) == ((int(act[0]), int(act[-1])) if act.size else (-1, -1)), i
        assert pk[i] == np.max(np.abs(seg)), i


def test_two_job_threads_get_their_own_results(env):
    """analyze(A), analyze(B), write(A) must not hand A the plan of B: every call of two concurrent threads equals
    the serial result."""
    torch, z, idx = env
    from vocalie_tts_b200 import post
    jobs = []
    for name in ("balance", "pauses", "many", "short"):
        flat, seg_off = _job(z, name)
        chunks = [flat[seg_off[i]:seg_off[i + 1]].copy() for i in range(len(seg_off) - 1)]
        jobs.append((name, chunks))
    serial = {name: post._apply_inter_chunk_gap(chunks, sr=SR, gap_ms=250) for name, chunks in jobs}
    errors = []

    def hammer(tid):
        try:
            torch.cuda.set_device(0)
            for it in range(12):
                name, chunks = jobs[(it + tid) % len(jobs)]
                got = post._apply_inter_chunk_gap(chunks, sr=SR, gap_ms=250)
                if not np.array_equal(got.view(np.uint32), serial[name].view(np.uint32)):
                    errors.append((tid, it, name))
                s, e = post._find_active_range(chunks[0], threshold=0.002, min_silence_frames=480)
                if (s, e) != po.find_active_range(chunks[0], threshold=0.002, min_silence_frames=480):
                    errors.append((tid, it, name, "range"))
        except Exception as exc:  # noqa: BLE001
            errors.append((tid, repr(exc)))

    ts = [threading.Thread(target=hammer, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:4]


def test_two_threads_apply_minimal_edit_files(env, tmp_path):
    torch, z, idx = env
    from vocalie_tts_b200 import post, wav
    names = ["balance", "pauses", "many", "single"]
    for n in names:
        wav.write_pcm16(tmp_path / f"{n}.wav", z[f"raw_{n}_250"], SR)
    # the goldens are read BEFORE the threads start: NpzFile decompresses lazily through one zipfile handle, which two
    # threads cannot share (a run on the GPU box failed with BadZipFile("Overlapped entries") inside the test itself)
    want = {n: (np.asarray(z[f"ame_tn_{n}_250"]), json.loads(bytes(z[f"ame_tn_meta_{n}_250"]).decode())) for n in names}
    errors = []

    def worker(tid):
        try:
            for it in range(8):
                n = names[(it + 2 * tid) % len(names)]
                out = tmp_path / f"{n}_{tid}_{it}.wav"
                res = post.apply_minimal_edit(tmp_path / f"{n}.wav", out, trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0)
                q, _ = wav.read_pcm16(out)
                golden, meta = want[n]
                if not np.array_equal(q, golden) or res["peak_before"] != meta["peak_before"]:
                    errors.append((tid, it, n))
        except Exception as exc:  # noqa: BLE001
            errors.append((tid, repr(exc)))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:4]
