"""Chunk-level data parallelism for one job on N GPUs of a node.

The chunks of a job are independent through HiFT and through the stitch pass, so they are sharded across
ranks with no data-path collective.  What is exchanged:

  * reference order (``granularity="job"``): the whole-file edit needs the file's trim range and its ONE
    peak (apply_minimal_edit on the stitched file, backend/services/tts_service.py:195-207,
    backend/shared/audio_edit.py:44-66).  Every rank reduces its own part of the raw file to three
    scalars (first active sample, last active sample, max|x|) and one ``all_reduce(MAX)`` of an int64[3]
    merges them; the min-silence rule of ``_find_active_range`` is applied once to the global range.
  * output assembly: every rank's pieces go STRAIGHT into their final position ``out[dst:dst+n]`` on rank 0
    with grouped send / recv (``ncclSend`` / ``ncclRecv`` over NVLink on GPUs, gloo in the CPU tests) - no
    staging buffer, no second copy.  Piece geometry is known on every rank: raw chunk lengths are ``480 T``
    (static), and data-dependent lengths (per-chunk trim) are all-gathered first (a few hundred int64).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

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


Piece = Tuple[int, int, int]   # (src offset in the rank's local buffer, samples, dst offset in the final file)


def stitched_pieces(all_lens: np.ndarray, gap: int, shards: Sequence[Sequence[int]]) -> List[List[Piece]]:
    """Pieces of a stitched job: rank r holds ``[chunk][gap][chunk][gap]...`` for its chunks (every chunk but the
    job's last followed by its gap); a run of globally consecutive chunks is contiguous on both sides."""
    lens = np.asarray(all_lens, dtype=np.int64)
    n_total = lens.size
    offs = global_offsets(lens, gap)
    out: List[List[Piece]] = []
    for ids in shards:
        pieces: List[Piece] = []
        if ids:
            l = lens[np.asarray(ids)]
            local_off = np.concatenate([[0], np.cumsum(l + gap)[:-1]])
            for first, count in _runs(ids):
                last = first + count - 1
                src0 = int(local_off[first])
                n = int(local_off[last] + l[last] - src0)
                if ids[last] != n_total - 1:          # the gap after the run belongs to the file
                    n += int(gap)
                pieces.append((src0, n, int(offs[ids[first]])))
        out.append(pieces)
    return out


def assemble_pieces(local, pieces_by_rank: Sequence[Sequence[Piece]], total: int, *, dst: int = 0, group=None, out=None):
    """Move every rank's pieces straight into ``out[dst_off:dst_off+n]`` on rank ``dst``: one grouped batch of
    send / recv operations whose receive buffers ARE the destination slices.  Returns ``out[:total]`` on ``dst``,
    None elsewhere."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    ops = []
    # torch's NCCL binding has no int16: PCM_16 pieces travel as bytes (same memory, no copy)
    wire = (lambda t: t.view(torch.uint8)) if local.dtype == torch.int16 else (lambda t: t)
    if rank == dst:
        if out is None:
            out = torch.empty(max(total, 1), dtype=local.dtype, device=local.device)
        for r, pieces in enumerate(pieces_by_rank):
            for src0, n, d0 in pieces:
                if n <= 0:
                    continue
                if r == rank:
                    out[d0:d0 + n].copy_(local[src0:src0 + n])
                else:
                    ops.append(dist.P2POp(dist.irecv, wire(out[d0:d0 + n]), dist.get_global_rank(group, r) if group is not None else r, group))
    else:
        for src0, n, d0 in pieces_by_rank[rank]:
            if n > 0:
                ops.append(dist.P2POp(dist.isend, wire(local[src0:src0 + n]), dist.get_global_rank(group, dst) if group is not None else dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out[:total] if rank == dst else None


def assemble_on_rank0(local_audio, local_lens: Sequence[int], chunk_ids: Sequence[int], n_total: int, gap: int,
                      shards: Sequence[Sequence[int]], *, group=None, out=None):
    """Assembly for data-dependent chunk lengths (``granularity="chunk"``: every chunk was trimmed on its own).

    ``local_audio``: this rank's stitched shard (1-D tensor, CPU or CUDA), chunk j at
    ``sum_{j'<j}(len_j' + gap)``; ``local_lens``: its per-chunk output lengths; ``shards``: the chunk
    ids of every rank (same on all ranks).  The lengths are all-gathered, then the pieces are sent straight into
    place.  Returns (final audio tensor on rank 0 | None, total samples).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = local_audio.device
    max_n = max(len(s) for s in shards)
    lens_pad = torch.zeros(max_n, dtype=torch.int64, device=dev)
    if len(local_lens):
        lens_pad[: len(local_lens)] = torch.as_tensor(np.asarray(local_lens, dtype=np.int64), device=dev)
    all_pad = torch.zeros(world * max_n, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_pad, lens_pad, group=group) if dev.type == "cuda" else \
        dist.all_gather(list(all_pad.view(world, max_n).unbind(0)), lens_pad, group=group)
    all_host = all_pad.view(world, max_n).cpu().numpy()
    all_lens = np.zeros(n_total, dtype=np.int64)
    for r, ids in enumerate(shards):
        if ids:
            all_lens[np.asarray(ids)] = all_host[r, : len(ids)]
    total = final_length(all_lens, gap)
    pieces = stitched_pieces(all_lens, gap, shards)
    # the trailing gap of the file's last piece may have been cut by the producer
    pieces = [[(s, min(n, total - d), d) for s, n, d in p] for p in pieces]
    out = assemble_pieces(local_audio, pieces, total, group=group, out=out)
    return out, total


class SharedHostBuffer:
    """Host memory mapped by every rank of the node (a file under /dev/shm) and page-locked in each process
    (cudaHostRegister): every rank copies ITS pieces of the finished file device -> host over its own PCIe link,
    straight into place - no gather to rank 0, no single-link read-back of the whole job.  Rank 0 creates the file,
    the others map it after a barrier; ``close`` unregisters and (rank 0) unlinks."""

    def __init__(self, name: str, numel: int, dtype, *, group=None):
        import os
        import torch
        import torch.distributed as dist
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.path = os.path.join("/dev/shm", name)
        self.numel, self.dtype = int(numel), dtype
        nbytes = self.numel * torch.empty(0, dtype=dtype).element_size()
        if self.rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        if dist.is_initialized():
            dist.barrier(group)
        self.tensor = torch.from_file(self.path, shared=True, size=self.numel, dtype=dtype)
        # completion flags, one int64 per rank, in the same kind of mapping (plain host stores: no GPU synchronisation):
        # flags[r] = index of the last job whose pieces rank r has finished copying
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.rank == 0:
            with open(self.path + ".flags", "wb") as f:
                f.write((-1).to_bytes(8, "little", signed=True) * self.world)
        if dist.is_initialized():
            dist.barrier(group)
        self.flags = torch.from_file(self.path + ".flags", shared=True, size=self.world, dtype=torch.int64)
        self._registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed with status {int(rc)}")
            self._registered = True

    def publish(self, job_index: int) -> None:
        """This rank's pieces of job ``job_index`` are in the buffer (call after ShardedJob.wait_host())."""
        self.flags[self.rank] = int(job_index)

    def wait_complete(self, job_index: int, timeout_s: float = 60.0) -> None:
        """Block (host-side polling of the shared flags) until EVERY rank has published ``job_index``."""
        import time
        t0 = time.monotonic()
        while int(self.flags.min()) < int(job_index):
            if time.monotonic() - t0 > timeout_s:
                raise TimeoutError(f"host assembly of job {job_index} incomplete: flags {self.flags.tolist()}")
            time.sleep(0.0002)

    def close(self):
        import os
        import torch
        import torch.distributed as dist
        if self._registered:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        self.tensor = None
        self.flags = None
        if dist.is_initialized():
            dist.barrier(self.group)
        if self.rank == 0:
            for pth in (self.path, self.path + ".flags"):
                try:
                    os.unlink(pth)
                except OSError:
                    pass


# ------------------------------------------------------------------------------------ reference order, sharded
def merge_file_range(first_g: int, last_g: int, n_file: int, *, trim: bool, min_silence_frames: int) -> Tuple[int, int]:
    """``_find_active_range`` on the whole file from the merged extremes (tts_pipeline.py:192-209) followed by
    apply_minimal_edit's guard (audio_edit.py:53: ``0 <= start < end <= len`` or leave the file alone)."""
    if not trim or n_file == 0 or last_g < 0:
        return 0, n_file
    start, end = int(first_g), int(last_g) + 1
    if start < min_silence_frames:
        start = 0
    if n_file - end < min_silence_frames:
        end = n_file
    if not (0 <= start < end <= n_file):
        return 0, n_file
    return start, end


@dataclass
class ShardedJobResult:
    audio: "object"            # final file on rank 0 (device tensor, int16 or float32), None elsewhere
    total_samples: int
    raw_samples: int
    edit: Optional[dict]


class ShardedJob:
    """One job in reference order over the ranks of a process group: every rank vocodes and stitches its chunks,
    one int64[3] all-reduce gives the file's trim range and peak, every rank edits its own part of the file and the
    parts are sent straight into place on rank 0."""

    def __init__(self, pipe, T_all: Sequence[int], shards: Sequence[Sequence[int]], *, group=None):
        import torch.distributed as dist
        from . import post as _post
        if pipe.granularity != "job":
            raise ValueError("ShardedJob runs the reference order: build the pipeline with granularity='job'")
        if pipe.editing and pipe.edit != "minimal_edit":
            raise ValueError("sharded jobs support edit='minimal_edit' (the variant run_tts_job calls) only")
        self.pipe, self.group = pipe, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.T_all = np.asarray(T_all, dtype=np.int64)
        self.shards = [list(s) for s in shards]
        self.n_total = int(self.T_all.size)
        pipe.set_shard(self.shards[self.rank], self.n_total)
        o = pipe.opts
        gap_on = o["chunk_gap_ms"] > 0 and self.n_total > 1
        self.gap = _post._ms_to_frames(pipe.sr, o["chunk_gap_ms"]) if gap_on else 0
        if not gap_on:
            pipe.stitch_head = pipe.stitch_tail = 1          # plain concatenate: no fades anywhere (tts_pipeline.py:171-172)
            pipe.opts = dict(o, chunk_gap_ms=0)
        self.lens = self.T_all * pipe.spf
        self.n_raw = final_length(self.lens, self.gap)
        self.raw_pieces = stitched_pieces(self.lens, self.gap, self.shards)
        self.local_T = self.T_all[np.asarray(self.shards[self.rank], dtype=np.int64)].astype(np.int32)
        mine = self.raw_pieces[self.rank]
        self.runs_off = np.concatenate([[0], np.cumsum([n for _, n, _ in mine])]).astype(np.int64)
        self.min_sil = int(pipe.sr * (int(o["silence_min_ms"]) / 1000.0))

    def run_device(self, mel, *, f0=None, phase_vec=None, noise=None, seed: int = 0, out=None,
                   max_frames: Optional[int] = None, host_out=None) -> ShardedJobResult:
        """``mel``: this rank's chunks, float32 CUDA [sum(local_T), 80].  ``max_frames`` bounds the mel frames per
        vocoder call (length bucketing of long jobs, HiFTVocoder.forward_bucketed)."""
        import torch
        pipe = self.pipe
        T = self.local_T
        n = int(T.astype(np.int64).sum()) * pipe.spf
        with torch.cuda.device(pipe.voc.device):
            wav = pipe._buf("_wav", n + 4, torch.float32)
            if T.size:
                pipe.voc.forward_bucketed(mel, T, f0=f0, phase_vec=phase_vec, noise=noise, seed=seed, out=wav, max_frames=max_frames)
            seg_off = np.concatenate([[0], np.cumsum(T.astype(np.int64) * pipe.spf)])
            res = self.post_device(wav, seg_off, out=out, host_out=host_out)
            from . import post as _post
            pipe.last_launches = (pipe.voc.last_launches + _post.last_launch_count() - pipe.voc._last_call_launches) if T.size \
                else _post.last_launch_count()
        return res

    def post_device(self, wav, seg_off, *, out=None, ops=None, host_out=None) -> ShardedJobResult:
        """Stitch -> (all-reduce -> whole-file edit) -> assembly, from this rank's packed raw chunks.  ``ops`` are the
        four device passes (default: the CUDA kernels; the gloo CPU tests inject the numpy oracle to exercise this
        host logic without a GPU).  ``host_out`` (a 1-D host tensor every rank can write, e.g. SharedHostBuffer.tensor):
        the file is assembled on the HOST instead - every rank copies its own pieces device -> host into place on a side
        stream (asynchronously: ``wait_host()`` before reading) and nothing is gathered to rank 0."""
        import torch
        import torch.distributed as dist
        pipe = self.pipe
        o = pipe.opts
        dev = wav.device
        ops = ops or _CudaOps(pipe)
        seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        n_chunks = seg_off.size - 1
        n_local_raw = int(self.runs_off[-1])
        if not pipe.editing:
            self._before_overwrite()
            local = ops.stitch(wav, seg_off, n_local_raw, final=True) if n_chunks else torch.zeros(0, device=dev)
            if host_out is not None:
                self._to_host(local, self.raw_pieces[self.rank], host_out)
                return ShardedJobResult(None, self.n_raw, self.n_raw, None)
            final = assemble_pieces(local, self.raw_pieces, self.n_raw, group=self.group, out=out)
            return ShardedJobResult(final, self.n_raw, self.n_raw, None)
        stats = torch.tensor([-(1 << 62), -1, 0], dtype=torch.int64, device=dev)     # [-first, last, peak bits]
        x = None
        if n_chunks:
            raw = ops.stitch(wav, seg_off, n_local_raw, final=False)        # PCM_16: stage B reads the file back
            x = ops.decode(raw, n_local_raw)
            fl, pk = ops.stats(x, self.runs_off, o["silence_threshold"])
            goff = torch.as_tensor(np.asarray([d for _, _, d in self.raw_pieces[self.rank]], dtype=np.int64), device=dev)
            active = fl[:, 1] >= 0
            first = torch.where(active, fl[:, 0] + goff, torch.full_like(goff, 1 << 62)).min()
            last = torch.where(active, fl[:, 1] + goff, torch.full_like(goff, -1)).max()
            stats = torch.stack([-first, last, pk.max().view(torch.int32).to(torch.int64)])   # non-negative floats order like their bits
        dist.all_reduce(stats, op=dist.ReduceOp.MAX, group=self.group)
        neg_first, last_g, peak_bits = (int(v) for v in stats.cpu().tolist())
        peak = float(np.array([peak_bits], dtype=np.int64).astype(np.int32).view(np.float32)[0])
        start_g, end_g = merge_file_range(-neg_first, last_g, self.n_raw, trim=o["trim_silence"], min_silence_frames=self.min_sil)
        total = end_g - start_g
        # every rank's pieces of the edited file: its runs clipped to [start_g, end_g)
        pieces: List[List[Piece]] = []
        for r in range(self.world):
            src, pr = 0, []
            for _, nrun, d0 in self.raw_pieces[r]:
                s = min(max(start_g - d0, 0), nrun)
                e = min(max(end_g - d0, s), nrun)
                pr.append((src, e - s, d0 + s - start_g))
                src += e - s
            pieces.append(pr)
        local = torch.zeros(0, device=dev)
        if n_chunks:
            rng = np.zeros((self.runs_off.size - 1, 2), dtype=np.int64)
            for i, (_, nrun, d0) in enumerate(self.raw_pieces[self.rank]):
                s = min(max(start_g - d0, 0), nrun)
                rng[i] = (s, min(max(end_g - d0, s), nrun))
            self._before_overwrite()
            local = ops.edit(x, self.runs_off, rng, peak if o["normalize"] else 0.0)
        if host_out is not None:
            self._to_host(local, pieces[self.rank], host_out)
            final = None
        else:
            final = assemble_pieces(local, pieces, total, group=self.group, out=out)
        target_peak = 10 ** (o["target_dbfs"] / 20.0)
        normalized = bool(o["normalize"] and peak > 0.0 and target_peak > 0.0)
        edit = {"start_sample": start_g, "end_sample": end_g, "peak_before": peak,
                "gain": (target_peak / peak) if normalized else 1.0,
                "trimmed": bool(o["trim_silence"] and 0 <= start_g < end_g <= self.n_raw), "normalized": normalized,
                "target_dbfs": o["target_dbfs"], "edit": pipe.edit}
        return ShardedJobResult(final, total, self.n_raw, edit)


    # ---- host assembly: this rank's pieces -> their place in a host buffer shared by the ranks, on a side stream
    def _before_overwrite(self):
        """The previous job's device->host copies read the buffer the next pass is about to overwrite."""
        import torch
        ev = getattr(self, "_copied", None)
        if ev is not None and torch.cuda.is_available():
            torch.cuda.current_stream().wait_event(ev)

    def _to_host(self, local, pieces, host_out):
        import torch
        if not local.is_cuda:                                   # CPU tests
            for src0, n, d0 in pieces:
                if n > 0:
                    host_out[d0:d0 + n].copy_(local[src0:src0 + n])
            return
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=local.device)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            for src0, n, d0 in pieces:
                if n > 0:
                    host_out[d0:d0 + n].copy_(local[src0:src0 + n], non_blocking=True)
            self._copied_prev = getattr(self, "_copied", None)
            self._copied = torch.cuda.Event()
            self._copied.record(self._copy_stream)

    def wait_host(self, previous: bool = False):
        """Block until this rank's pieces of the last (``previous``: the one before the last) host-assembled job have
        landed; then ``SharedHostBuffer.publish`` / ``wait_complete`` tell the ranks apart."""
        ev = getattr(self, "_copied_prev" if previous else "_copied", None)
        if ev is not None:
            ev.synchronize()


class _CudaOps:
    """The four device passes of a sharded job on the CUDA kernels (csrc/vt_post.cu)."""

    def __init__(self, pipe):
        self.pipe = pipe

    def stitch(self, wav, seg_off, n_local_raw, *, final):
        import torch
        from . import post as _post
        pipe = self.pipe
        pcm = True if not final else pipe.opts["out_pcm16"]
        prm = pipe.stitch_params(len(seg_off) - 1, out_pcm16=pcm)
        buf = pipe._buf("_out" if final else "_raw", n_local_raw + int(prm.gap_frames), torch.int16 if pcm else torch.float32)
        _post.post_process_device(wav, seg_off, prm, out=buf, read_back=False)
        return buf

    def decode(self, raw, n):
        import torch
        from . import _lib
        x = self.pipe._buf("_rawf", n + 4, torch.float32)
        _lib.check(_lib.load_library().vt_pcm16_decode(int(raw.data_ptr()), int(x.data_ptr()), n,
                                                       int(torch.cuda.current_stream().cuda_stream)), "vt_pcm16_decode")
        return x

    def stats(self, x, runs_off, threshold):
        from . import post as _post
        return _post.stats_device(x, runs_off, threshold=threshold)

    def edit(self, x, runs_off, rng, peak):
        import torch
        from . import post as _post
        pipe = self.pipe
        out = pipe._buf("_out", int(runs_off[-1]), torch.int16 if pipe.opts["out_pcm16"] else torch.float32)
        _post.post_process_device(x, runs_off, pipe.edit_params(), out=out, read_back=False,
                                  range_override=torch.as_tensor(rng, device=x.device),
                                  peak_override=torch.tensor([peak], dtype=torch.float32, device=x.device))
        return out
