"""Chunk-level data parallelism for one job on N GPUs of a node.

The chunks of a job are independent through HiFT and through per-chunk post-processing, so they are
sharded across ranks with no data-path collective.  The one exchange step is output assembly:
every rank holds ``[chunk][gap][chunk][gap]...`` for its chunks; rank 0 needs them interleaved in
job order.  Trimmed lengths are data dependent, so the ranks first all-gather their per-chunk
output lengths (a few hundred int64), compute the global offsets (exclusive scan of
``len_i + gap``), then the stitched shards are gathered to rank 0 (NCCL over NVLink on GPUs, gloo
in the CPU tests) and copied run by run into place.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def contiguous_shards(n_chunks: int, world: int) -> List[List[int]]:
    """Equal contiguous ranges in job order (what bench.py uses for equal-length chunks)."""
    base, rem = divmod(n_chunks, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append(list(range(start, start + n)))
        start += n
    return out


def stitch_flags(chunk_ids: Sequence[int], n_total: int):
    """(stitch_head, stitch_tail) for a rank holding ``chunk_ids`` (sorted global indices): only the
    rank with the job's first chunk skips its fade-in, only the rank with the last chunk skips the
    fade-out and the trailing gap (reference _apply_inter_chunk_gap, tts_pipeline.py:179-187)."""
    ids = list(chunk_ids)
    return (1 if ids and ids[0] == 0 else 0), (1 if ids and ids[-1] == n_total - 1 else 0)


def global_offsets(all_lens: np.ndarray, gap: int) -> np.ndarray:
    """Exclusive scan of ``len_i + gap`` over the job's chunks -> offset of every chunk in the final file."""
    lens = np.asarray(all_lens, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(lens + int(gap))[:-1]]) if lens.size else np.zeros(0, np.int64)


def final_length(all_lens: np.ndarray, gap: int) -> int:
    lens = np.asarray(all_lens, dtype=np.int64)
    return int(lens.sum() + max(lens.size - 1, 0) * int(gap)) if lens.size else 0


def _runs(chunk_ids: Sequence[int]):
    """Maximal runs of globally consecutive chunks inside one rank's (sorted) list: [(first_pos, count)]."""
    runs, i, ids = [], 0, list(chunk_ids)
    while i < len(ids):
        j = i
        while j + 1 < len(ids) and ids[j + 1] == ids[j] + 1:
            j += 1
        runs.append((i, j - i + 1))
        i = j + 1
    return runs


def assemble_on_rank0(local_audio, local_lens: Sequence[int], chunk_ids: Sequence[int], n_total: int, gap: int,
                      shards: Sequence[Sequence[int]], *, group=None, out=None):
    """Gather the stitched shards to rank 0 and interleave them in job order.

    ``local_audio``: this rank's stitched shard (1-D tensor, CPU or CUDA), chunk j at
    ``sum_{j'<j}(len_j' + gap)``; ``local_lens``: its per-chunk output lengths; ``shards``: the chunk
    ids of every rank (same on all ranks).  Returns (final audio tensor on rank 0 | None, total samples).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local_audio.device
    max_n = max(len(s) for s in shards)
    lens_pad = torch.zeros(max_n, dtype=torch.int64, device=dev)
    if len(local_lens):
        lens_pad[: len(local_lens)] = torch.as_tensor(np.asarray(local_lens, dtype=np.int64), device=dev)
    all_pad = [torch.zeros_like(lens_pad) for _ in range(world)]
    dist.all_gather(all_pad, lens_pad, group=group)
    all_lens = np.zeros(n_total, dtype=np.int64)
    for r, ids in enumerate(shards):
        if ids:
            all_lens[np.asarray(ids)] = all_pad[r][: len(ids)].cpu().numpy()
    total = final_length(all_lens, gap)
    # fixed-size gather (shards are padded to the largest one; the pad is never copied)
    cap = 0
    for ids in shards:
        cap = max(cap, int(all_lens[np.asarray(ids, dtype=np.int64)].sum() + len(ids) * gap) if ids else 0)
    cap = max(cap, 1)
    send = local_audio
    if send.numel() < cap:
        send = torch.zeros(cap, dtype=local_audio.dtype, device=dev)
        send[: local_audio.numel()] = local_audio
    else:
        send = send[:cap].contiguous()
    bufs = [torch.empty(cap, dtype=local_audio.dtype, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(send, bufs, dst=0, group=group)
    if rank != 0:
        return None, total
    offs = global_offsets(all_lens, gap)
    if out is None:
        out = torch.empty(max(total, 1), dtype=local_audio.dtype, device=dev)
    for r, ids in enumerate(shards):
        if not ids:
            continue
        l = all_lens[np.asarray(ids)]
        local_off = np.concatenate([[0], np.cumsum(l + gap)[:-1]])
        for first, count in _runs(ids):
            last = first + count - 1
            src0 = int(local_off[first])
            n = int(local_off[last] + l[last] - src0)
            # the gap after the run's last chunk belongs to the file unless it is the job's last chunk
            if ids[last] != n_total - 1:
                n += gap
            dst0 = int(offs[ids[first]])
            n = min(n, total - dst0)
            out[dst0:dst0 + n].copy_(bufs[r][src0:src0 + n])
    return out[:total], total
