"""world_size-2 / 3 gloo tests of the reference-order sharded job (distributed.ShardedJob): per-rank stitch ->
PCM_16 raw file -> int64[3] all-reduce (first, last, peak) -> per-rank whole-file edit -> pieces sent straight
into place on rank 0.  The four device passes are replaced by the numpy oracle (``ops=``) so the host logic -
piece geometry, the merged trim range, the cross-rank peak, the grouped send/recv - runs here without a GPU; the
expected files are the REFERENCE's own outputs (tests/golden/job_golden.npz: _apply_inter_chunk_gap -> sf.write ->
apply_minimal_edit on the stitched file)."""
import json
import os
import socket
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import post_oracle as po

SR = 24000
GOLD = Path(__file__).resolve().parent / "golden" / "job_golden.npz"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeVoc:
    device = torch.device("cpu")


class _NumpyOps:
    """The four passes of distributed._CudaOps restated on the numpy oracle (CPU tensors in and out)."""

    def __init__(self, pipe):
        self.pipe = pipe

    def stitch(self, wav, seg_off, n_local_raw, *, final):
        p = self.pipe
        x = wav.numpy()
        chunks = [x[seg_off[i]:seg_off[i + 1]] for i in range(len(seg_off) - 1)]
        gap = po.ms_to_frames(SR, p.opts["chunk_gap_ms"])
        parts = []
        n = len(chunks)
        for i, c in enumerate(chunks):
            a = np.array(c, dtype=np.float32, copy=True)
            if gap > 0:
                if not (i == n - 1 and p.stitch_tail):
                    po.fade_out(a, 240)
                if not (i == 0 and p.stitch_head):
                    po.fade_in(a, 240)
            parts.append(a)
            if gap > 0 and not (i == n - 1 and p.stitch_tail):
                parts.append(np.zeros(gap, np.float32))
        y = np.concatenate(parts)
        assert y.size == n_local_raw
        if final and not p.opts["out_pcm16"]:
            return torch.from_numpy(y)
        return torch.from_numpy(po.pcm16_encode(y))

    def decode(self, raw, n):
        return torch.from_numpy(po.pcm16_decode(raw.numpy()[:n]))

    def stats(self, x, runs_off, threshold):
        x = x.numpy()
        fl, pk = [], []
        for i in range(len(runs_off) - 1):
            seg = x[runs_off[i]:runs_off[i + 1]]
            act = np.flatnonzero(np.abs(seg) > np.float32(threshold))
            fl.append((int(act[0]), int(act[-1])) if act.size else (-1, -1))
            pk.append(float(np.max(np.abs(seg))) if seg.size else 0.0)
        return torch.tensor(fl, dtype=torch.int64).reshape(-1, 2), torch.tensor(pk, dtype=torch.float32)

    def edit(self, x, runs_off, rng, peak):
        p = self.pipe
        x = x.numpy()
        out = []
        for i in range(len(runs_off) - 1):
            seg = x[runs_off[i] + rng[i, 0]: runs_off[i] + rng[i, 1]]
            if peak > 0.0:
                seg = seg * (float(10 ** (p.opts["target_dbfs"] / 20.0)) / float(np.float32(peak)))
            out.append(np.clip(seg, -1.0, 1.0).astype(np.float32))
        y = np.concatenate(out) if out else np.zeros(0, np.float32)
        return torch.from_numpy(po.pcm16_encode(y) if p.opts["out_pcm16"] else y)


def _worker(rank, world, port, job, gap, tag, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vocalie_tts_b200 import distributed as D
        from vocalie_tts_b200.backend import shard_chunks
        from vocalie_tts_b200.pipeline import VocoderPipeline
        z = np.load(GOLD)
        lens = z[f"n_{job}"]
        flat = z[f"in_{job}"]
        off = np.concatenate([[0], np.cumsum(lens)])
        chunks = [flat[off[i]:off[i + 1]] for i in range(len(lens))]
        n_total = len(chunks)
        shards = D.contiguous_shards(n_total, world) if mode == "contiguous" else shard_chunks(lens.tolist(), world)
        edits = {"tn": (True, True, -1.0), "t": (True, False, -3.0), "n": (False, True, -6.0), "raw": (False, False, -1.0)}
        trim, norm, db = edits[tag]
        pipe = VocoderPipeline(_FakeVoc(), chunk_gap_ms=gap, trim_silence=trim, normalize=norm, target_dbfs=db, out_pcm16=True)
        # ShardedJob derives lengths from mel frames (480 samples each); the golden chunks have arbitrary lengths, so
        # the geometry is injected directly
        job_ = D.ShardedJob.__new__(D.ShardedJob)
        job_.pipe, job_.group, job_.rank, job_.world = pipe, None, rank, world
        job_.shards, job_.n_total = [list(s) for s in shards], n_total
        pipe.set_shard(job_.shards[rank], n_total)
        gap_on = gap > 0 and n_total > 1
        job_.gap = po.ms_to_frames(SR, gap) if gap_on else 0
        if not gap_on:
            pipe.stitch_head = pipe.stitch_tail = 1
            pipe.opts = dict(pipe.opts, chunk_gap_ms=0)
        job_.lens = lens.astype(np.int64)
        job_.n_raw = D.final_length(job_.lens, job_.gap)
        job_.raw_pieces = D.stitched_pieces(job_.lens, job_.gap, job_.shards)
        job_.runs_off = np.concatenate([[0], np.cumsum([n for _, n, _ in job_.raw_pieces[rank]])]).astype(np.int64)
        job_.min_sil = 480
        mine = [chunks[i] for i in job_.shards[rank]]
        wav = torch.from_numpy(np.concatenate(mine) if mine else np.zeros(0, np.float32))
        seg_off = np.concatenate([[0], np.cumsum([c.size for c in mine])]).astype(np.int64)
        res = job_.post_device(wav, seg_off, ops=_NumpyOps(pipe))
        if rank == 0:
            want = z[f"raw_{job}_{gap}"] if tag == "raw" else z[f"ame_{tag}_{job}_{gap}"]
            got = res.audio.numpy()
            ok = got.size == want.size and np.array_equal(got, want)
            msg = "ok" if ok else f"mismatch: got {got.size} want {want.size}"
            if ok and tag != "raw":
                meta = json.loads(bytes(z[f"ame_{tag}_meta_{job}_{gap}"]).decode())
                e = res.edit
                if not (e["trimmed"] == meta["trimmed"] and e["normalized"] == meta["normalized"]
                        and e["peak_before"] == meta["peak_before"] and abs(e["gain"] - meta["gain"]) <= 1e-15 * abs(meta["gain"])):
                    msg = f"meta mismatch {e} vs {meta}"
            q.put((msg, job, gap, tag, mode, world))
        else:
            assert res.audio is None
    except Exception as exc:  # surface worker failures in the parent
        import traceback
        q.put((f"rank {rank}: {type(exc).__name__}: {exc}\n{traceback.format_exc()}", job, gap, tag, mode, world))
    finally:
        dist.destroy_process_group()


CASES = [(2, "contiguous", "balance", 250, "tn"), (2, "lpt", "pauses", 250, "tn"), (3, "lpt", "many", 250, "tn"),
         (2, "lpt", "many", 250, "t"), (3, "contiguous", "many", 250, "n"), (2, "contiguous", "silent_edges", 250, "tn"),
         (2, "contiguous", "all_silent", 250, "tn"), (3, "lpt", "short", 250, "tn"), (2, "lpt", "many", 0, "tn"),
         (3, "contiguous", "single", 250, "tn"), (2, "lpt", "many", 250, "raw")]


@pytest.mark.parametrize("world,mode,job,gap,tag", CASES)
def test_sharded_reference_order_job_equals_the_reference_file(world, mode, job, gap, tag):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, job, gap, tag, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
    msg = q.get(timeout=5)
    assert msg[0] == "ok", msg


def test_merge_file_range_rules():
    from vocalie_tts_b200.distributed import merge_file_range
    assert merge_file_range(1 << 62, -1, 1000, trim=True, min_silence_frames=480) == (0, 1000)      # all silent
    assert merge_file_range(100, 900, 1000, trim=True, min_silence_frames=480) == (0, 1000)         # both edges within 20 ms
    assert merge_file_range(500, 400, 2000, trim=True, min_silence_frames=480) == (0, 2000)         # degenerate -> untouched
    assert merge_file_range(500, 1400, 2000, trim=True, min_silence_frames=480) == (500, 1401)
    assert merge_file_range(500, 1600, 2000, trim=True, min_silence_frames=480) == (500, 2000)
    assert merge_file_range(500, 1400, 2000, trim=False, min_silence_frames=480) == (0, 2000)
