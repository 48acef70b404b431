"""world_size-2 (and 3) gloo tests of the multi-GPU host logic: shard -> per-rank stitch with
head/tail flags -> all-gather of lengths -> gather + interleave on rank 0 must equal the
single-process reference stitch (oracle of tts_pipeline._apply_inter_chunk_gap)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import post_oracle as po

SR = 24000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_stitch(chunks, head, tail, gap_ms):
    """Oracle of the kernel's stitch with stitch_head/stitch_tail flags for one rank's chunks."""
    gap = po.ms_to_frames(SR, gap_ms)
    parts, lens = [], []
    n = len(chunks)
    for i, c in enumerate(chunks):
        a = np.array(c, dtype=np.float32, copy=True)
        if not (i == n - 1 and tail):
            po.fade_out(a, 240)
        if not (i == 0 and head):
            po.fade_in(a, 240)
        parts.append(a)
        lens.append(a.size)
        if not (i == n - 1 and tail):
            parts.append(np.zeros(gap, np.float32))
    return (np.concatenate(parts) if parts else np.zeros(0, np.float32)), lens


def _worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vocalie_tts_b200 import distributed as D
        from vocalie_tts_b200.backend import shard_chunks
        rng = np.random.default_rng(7)
        n_total = 11
        lens = rng.integers(300, 4000, n_total)
        chunks = [(rng.standard_normal(int(n)) * 0.3).astype(np.float32) for n in lens]
        shards = D.contiguous_shards(n_total, world) if mode == "contiguous" else shard_chunks(lens.tolist(), world)
        ids = shards[rank]
        head, tail = D.stitch_flags(ids, n_total)
        local, local_lens = _shard_stitch([chunks[i] for i in ids], head, tail, 250)
        out, total = D.assemble_on_rank0(torch.from_numpy(local), local_lens, ids, n_total, po.ms_to_frames(SR, 250), shards)
        if rank == 0:
            want = po.apply_inter_chunk_gap(chunks, sr=SR, gap_ms=250)
            ok = (total == want.size) and np.array_equal(out.numpy().view(np.uint32), want.view(np.uint32))
            q.put(("ok" if ok else f"mismatch total={total} want={want.size}", mode, world))
        else:
            assert out is None
    except Exception as exc:  # surface worker failures in the parent
        q.put((f"rank {rank}: {type(exc).__name__}: {exc}", mode, world))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "contiguous"), (2, "lpt"), (3, "lpt")])
def test_sharded_job_assembles_to_single_process_result(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
    msg = q.get(timeout=5)
    assert msg[0] == "ok", msg


def test_shard_helpers():
    from vocalie_tts_b200 import distributed as D
    assert D.contiguous_shards(10, 4) == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert D.contiguous_shards(2, 4) == [[0], [1], [], []]
    assert D.stitch_flags([0, 1], 4) == (1, 0) and D.stitch_flags([2, 3], 4) == (0, 1) and D.stitch_flags([], 4) == (0, 0)
    assert D.stitch_flags([0, 3], 4) == (1, 1)
    assert list(D.global_offsets([10, 20, 5], 3)) == [0, 13, 36] and D.final_length([10, 20, 5], 3) == 41
    assert D._runs([0, 1, 2, 5, 7, 8]) == [(0, 3), (3, 1), (4, 2)]
