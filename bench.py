#!/usr/bin/env python
"""Headline benchmark: audio-seconds per second of the HiFT vocoder + post-processing path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--operand fp16|bf16|fp32]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): every rank turns
64 chunks x 10 s of synthetic log-mel (T = 500 frames, 80 bins) into 24 kHz audio with random-init
HiFT weights, then runs the per-chunk post-processing (silence trim with zero-cross snap, edge
fades, -1 dBFS peak normalise, 250 ms gap stitching).  With N > 1 (torchrun, one rank per GPU) the
chunks of the job are sharded 64 per rank (weak scaling; N = 8 is configs[2], 512 chunks) and the
stitched shards are gathered to rank 0 over NCCL - the only exchange step of the path.

One JSON line is printed by rank 0; see the task contract for the keys.  `value` = whole-job
audio seconds / device time with the mels already in HBM; `e2e` = the same through
VocoderPipeline.submit() / collect() with pinned HOST mels in and host audio out (job i+1 is enqueued while job
i's audio crosses PCIe on a second stream; every copy is inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CHUNKS_PER_RANK = 64
T_FRAMES = 500
SR = 24000
SPF = 480
GAP_MS = 250


def synth_mel_host(n_chunks, T, seed):
    """clamp(N(-5, 2^2), ln(1e-5), 2) rounded to bf16, frame-major [sum_T, 80] float32."""
    import torch
    g = torch.Generator().manual_seed(seed)
    m = torch.randn(n_chunks * T, 80, generator=g) * 2.0 - 5.0
    return m.clamp_(float(np.log(1e-5)), 2.0).to(torch.bfloat16).to(torch.float32)


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples in the upper half of the observed power range
        thr = (max(pw) + min(pw)) / 2 if pw else 0
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_sample(n_chunks, T, threads):
    """The reference's CPU implementation of the path on a bounded sample: torch fp32 restatement of
    upstream HiFT (oracle/hift_oracle.py, chatterbox-tts is not installable offline) + the numpy
    post-processing restatement (oracle/post_oracle.py), chunk by chunk as run_tts_pipeline does.
    Returns (audio_seconds, wall_seconds)."""
    import torch
    from oracle import hift_oracle as H
    from oracle import post_oracle as po
    torch.set_num_threads(threads)
    W = H.fold_weight_norm(H.make_state_dict(0, "init"))
    mels = [H.synth_mel(T, 2, b) for b in range(n_chunks)]
    pn = [H.synth_noise(T, 2, b) for b in range(n_chunks)]
    t0 = time.perf_counter()
    chunks = []
    for b in range(n_chunks):
        wav = H.hift_inference(mels[b], W, f0=None, phase_vec=pn[b][0], noise=pn[b][1]).numpy()
        y, _ = po.minimal_post_process_array(wav, SR)
        chunks.append(y)
    out = po.apply_inter_chunk_gap(chunks, sr=SR, gap_ms=GAP_MS)
    dt = time.perf_counter() - t0
    assert out.size > 0
    return n_chunks * T * SPF / SR, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_chunks = 2
    # warm-up + K bounded steps (each 2 chunks x 10 s = 20 s of audio)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(1, 100, threads)
    times = []
    audio = 0.0
    for _ in range(args.steps):
        a, dt = cpu_reference_sample(n_chunks, T_FRAMES, threads)
        audio = a
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = audio / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "audio-sec/sec (HiFT vocoder+post)", "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Chatterbox HiFT batch of {CHUNKS_PER_RANK} chunks x 10 s mel + trim/normalise/gap post per GPU "
                               f"(CPU arm: bounded sample of {n_chunks} chunks x 10 s per step)"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": f"{n_chunks} chunks x T={T_FRAMES} per step, torch {threads} threads + numpy post"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from vocalie_tts_b200.hift import HiFTVocoder, random_state_dict, algorithmic_flops_per_frame
    from vocalie_tts_b200.pipeline import VocoderPipeline

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    voc = HiFTVocoder(random_state_dict(0), operand=args.operand)
    voc.set_profiling(True)
    pipe = VocoderPipeline(voc, chunk_gap_ms=GAP_MS)
    n_chunks, T = args.chunks, args.frames
    Ts = np.full(n_chunks, T, dtype=np.int32)
    mel_host = pipe.pinned_input(n_chunks * T)
    mel_host.copy_(synth_mel_host(n_chunks, T, 1000 + rank))
    mel_dev = mel_host.to(dev)
    audio_s_rank = n_chunks * T * SPF / SR
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    from vocalie_tts_b200 import distributed as D
    gap_frames = GAP_MS * SR // 1000
    shards = D.contiguous_shards(n_chunks * world, world)
    pipe.set_shard(shards[rank], n_chunks * world)
    shard_cap = n_chunks * T * SPF + n_chunks * gap_frames
    final = [torch.empty(world * shard_cap, dtype=torch.float32, device=dev) if (world > 1 and rank == 0) else None]
    host_final = torch.empty(world * shard_cap, dtype=torch.float32).pin_memory() if (world > 1 and rank == 0) else None
    totals = [0]

    def step_device(seed):
        res = pipe.run_device(mel_dev, Ts, seed=seed, read_back=world > 1)
        if world > 1:
            # output assembly: all-gather of per-chunk lengths + NCCL gather of the stitched shards to rank 0
            lens = res.segments[:, 5].astype(np.int64)
            out, total = D.assemble_on_rank0(res.audio[: res.total_samples], lens, shards[rank], n_chunks * world,
                                             gap_frames, shards, out=final[0])
            totals[0] = total
        return res

    # End to end = the serving loop a caller runs: job i+1 is enqueued while job i's audio is still crossing PCIe
    # (two output slots, copies on a second stream).  Every step's host->device and device->host copies are inside
    # the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    if world > 1 and rank == 0:
        finals = [final[0], torch.empty_like(final[0])]
        host_finals = [host_final, torch.empty_like(host_final).pin_memory()]
        copied = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(steps, seed0):
        """Host mels in, host audio out, `steps` jobs back to back.  One GPU: the public VocoderPipeline.submit() /
        collect().  Several GPUs: the same copies around the sharded step; rank 0 reads every assembled job back."""
        d2h = 0
        if world == 1:
            pending = None
            for i in range(steps):
                t = pipe.submit(mel_host, Ts, seed=seed0 + i)
                if pending is not None:
                    pipe.collect(pending)
                pending = t
            pipe.collect(pending)
            return pipe.last_d2h_bytes
        cur = torch.cuda.current_stream()
        for i in range(steps):
            slot = i & 1
            if rank == 0:
                cur.wait_event(copied[slot])          # the slot's previous job has left the device
                final[0] = finals[slot]
            mel_dev.copy_(mel_host, non_blocking=True)
            step_device(seed0 + i)
            if rank == 0:
                ready = torch.cuda.Event()
                ready.record(cur)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ready)
                    host_finals[slot][: totals[0]].copy_(finals[slot][: totals[0]], non_blocking=True)
                    copied[slot].record(copy_stream)
                d2h = totals[0] * 4
        copy_stream.synchronize()
        cur.synchronize()
        if rank == 0:
            final[0] = finals[0]
        return d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    prof = {"total_ms": 0.0, "resblock_ms": 0.0, "resblock_flops": 0.0, "resblock_launches": 0}
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 1)                    # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        step_device(100 + i)
        ev[i][1].record()
        launches += pipe.last_launches
        p = voc.read_profile()                # waits for this step's forward; the post kernels are still timed by ev
        for k in prof:
            prof[k] += p[k]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_local = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([ms_local], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = audio_s_rank * world / (ms * 1e-3)

    # ---- end to end through the public API: pinned host mels in, host audio out, every step
    e2e_loop(2, 7)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    e2e_steps = args.steps
    wall0 = time.perf_counter()
    t0.record()
    d2h = e2e_loop(e2e_steps, 200)
    t1.record()
    barrier()
    wall = (time.perf_counter() - wall0) / e2e_steps
    e2e_ms = max(t0.elapsed_time(t1) / e2e_steps, wall * 1e3)
    te = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = audio_s_rank * world / (float(te.item()) * 1e-3)

    if rank != 0:
        return
    hbm, tf, how = peaks()
    rb_ms = prof["resblock_ms"] / args.steps
    rb_tflops = prof["resblock_flops"] / args.steps / (rb_ms * 1e-3) / 1e12 if rb_ms > 0 else 0.0
    fwd_ms = prof["total_ms"] / args.steps
    line = {
        "metric": "audio-sec/sec (HiFT vocoder+post)", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.operand] + " operands, f32 accumulate",
        "data": "synthetic",
        "config": {"workload": f"Chatterbox HiFT batch of {n_chunks} chunks x {T * SPF / SR:g} s mel + trim/normalise/gap post per GPU",
                   "chunks_per_gpu": n_chunks, "mel_frames": T, "chunk_gap_ms": GAP_MS, "weights": "random-init (upstream init)",
                   "f0": "predicted (ConvRNNF0Predictor)", "noise": "in-kernel Philox",
                   "l2": "flushed between timed steps (256 MB write); per-step activations >> L2",
                   "exchange": "NCCL gather of stitched shards to rank 0" if world > 1 else "none"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(mel_host.numel() * 4),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": f"ResBlock convolutions ({prof['resblock_launches'] // args.steps} launches per forward: fused pairs at C=64/128, single convs at C=256)",
                     "achieved": rb_tflops, "peak": tf, "unit": "TFLOP/s", "frac": rb_tflops / tf,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the top kernel (k_pair_tc<64>, k=7,
                     # 3.84 M steps) from the ncu --set full capture summarised in profiles/r01_pair_c64_ncu.txt;
                     # algorithmic = 512 B per step = 1.966e9 B
                     "traffic": 1.916e9 if (n_chunks, T) == (CHUNKS_PER_RANK, T_FRAMES) else None,
                     "peak_source": how, "kernel_ms_per_step": rb_ms, "forward_ms_per_step": fwd_ms,
                     "kernel_share_of_step": rb_ms / ms if ms > 0 else None,
                     "path_tflops": algorithmic_flops_per_frame() * n_chunks * T / (ms * 1e-3) / 1e12},
    }
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        a, dt = cpu_reference_sample(1, 250, threads)      # warm-up (configs[0] shape)
        a, dt = cpu_reference_sample(4, T_FRAMES, threads)
        line["cpu_baseline"] = {"value": a / dt, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                "sample": f"4 chunks x T={T_FRAMES} (40 s audio): torch fp32 HiFT restatement on {threads} threads + numpy post"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--operand", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--chunks", type=int, default=CHUNKS_PER_RANK)
    ap.add_argument("--frames", type=int, default=T_FRAMES)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
