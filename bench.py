#!/usr/bin/env python
"""Headline benchmark: audio-seconds per second of the HiFT vocoder + post-processing path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_eager]
                    [--workload cfg2|cfg4] [--operand fp16|bf16|fp32]

Default workload = BASELINE.json configs[1] (the configuration the metric is quoted on): every rank turns 64 chunks x
10 s of synthetic log-mel (T = 500 frames, 80 bins) into 24 kHz audio with random-init HiFT weights, then the job
goes through post-processing IN THE REFERENCE'S ORDER: 250 ms gap stitch of the raw chunks -> PCM_16 raw file -> one
whole-file apply_minimal_edit (trim, ONE -1 dBFS peak, clip) -> PCM_16 (backend/shared/tts_pipeline.py:395-409,
backend/services/tts_service.py:195-207).  With N > 1 (torchrun, one rank per GPU) the job has 64 N chunks sharded 64
per rank (weak scaling; N = 8 is configs[2], 512 chunks): one int64[3] all-reduce merges the file's trim range and
peak, and every rank's part of the finished file is sent straight into place on rank 0 (grouped ncclSend/ncclRecv).

`--workload cfg4` = configs[3]: 2 048 chunks of T = 50 U{1..20} frames (seed 1004), LPT-sharded over the ranks by mel
length, every rank vocoding its chunks in length-bounded buckets; reports per-rank times and the max/mean imbalance.

One JSON line is printed by rank 0; see the task contract for the keys.  `value` = whole-job audio seconds / device time
with the mels already in HBM; `e2e` = the same through the public host API (pinned HOST mels in, HOST PCM_16 audio out,
every copy inside the timed region; one GPU: VocoderPipeline.submit()/collect(), job i+1 enqueued while job i's audio
crosses PCIe).  Extra objects of the N = 1 line: `roofline` (tensor: the 72 ResBlock convolutions), `roofline_post` (HBM:
configs[4], a 1 GiB post-only sweep), `latency_cfg1` (configs[0]: one 5 s chunk, host to host), `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import platform
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
CFG4_CHUNKS = 2048
CFG4_MAX_FRAMES = 32000          # mel frames per vocoder call (= the cfg2 batch: ~14 GB of workspace)
METRIC = "audio-sec/sec (HiFT vocoder+post)"


def synth_mel_host(n_frames, seed):
    """clamp(N(-5, 2^2), ln(1e-5), 2) rounded to bf16, frame-major [n_frames, 80] float32."""
    import torch
    g = torch.Generator().manual_seed(seed)
    m = torch.randn(n_frames, 80, generator=g) * 2.0 - 5.0
    return m.clamp_(float(np.log(1e-5)), 2.0).to(torch.bfloat16).to(torch.float32)


def cfg4_lengths():
    """BASELINE configs[3] / SURVEY 8(d): 2 048 chunks, T_b = 50 * U{1..20} (seed 1004)."""
    import torch
    g = torch.Generator().manual_seed(1004)
    return (50 * torch.randint(1, 21, (CFG4_CHUNKS,), generator=g)).numpy().astype(np.int64)


def cpu_model() -> str:
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (profiling recipe's clocks line): through NVML
    every 5 ms when pynvml is importable (the timed region of the default run is ~140 ms: nvidia-smi's 100 ms loop sees it
    once or twice), else the recipe's nvidia-smi loop."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.p = None
        self.nv = None          # (pynvml module, device handle) when NVML is usable: ~200 samples per second instead of 10
        self.extra = set()      # throttle reasons beyond the four of the profiling recipe's query (NVML sampler)
        self.first = 0          # first sample of the timed region (mark())
        self._stop = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:                    # CUDA ordinal -> NVML device through the PCI address (CUDA_VISIBLE_DEVICES may renumber)
            import torch
            pr = torch.cuda.get_device_properties(self.idx)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:       # noqa: BLE001  (older torch without the PCI fields)
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.idx)

    def _poll_nvml(self):
        nv, h = self.nv
        R = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
             "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        # every other reason NVML knows (GPU idle excepted) is reported under its own name: a clock below the maximum
        # should never appear without the reason the driver gives for it
        other = {"hw_power_brake_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0),
                 "sync_boost": getattr(nv, "nvmlClocksThrottleReasonSyncBoost", 0),
                 "applications_clocks_setting": getattr(nv, "nvmlClocksThrottleReasonApplicationsClocksSetting", 0),
                 "display_clock_setting": getattr(nv, "nvmlClocksThrottleReasonDisplayClockSetting", 0)}
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self._stop.is_set():
            try:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(self.idx), str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx),
                                  str(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)] +
                                 ["Active" if mask & bit else "Not Active" for bit in R.values()])
                for name, bit in other.items():
                    if bit and mask & bit:
                        self.extra.add(name)
            except Exception:   # noqa: BLE001
                pass
            self._stop.wait(0.005)

    def start(self):
        try:
            self.nv = self._nvml_handle()
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:       # noqa: BLE001  (no pynvml / NVML refused: the nvidia-smi loop of the profiling recipe)
            self.nv = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def mark(self):
        """The timed region starts here: clocks and power are reported from the samples taken after this call."""
        self.first = len(self.rows)

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv:
            self._stop.set()
            self.t.join(timeout=2)
        elif not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        out = self.summarize(self.rows, self.first, self.extra)
        out["sampler"] = "nvml" if self.nv else "nvidia-smi"
        return out

    @staticmethod
    def summarize(rows, first, extra=()):
        """rows: [index, sm MHz, max sm MHz, power W, hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap]
        as strings (the nvidia-smi query order); clocks and power from rows[first:], throttle reasons from every row."""
        sm, mx, pw, reasons = [], [], [], set()
        timed = rows[first:] if len(rows) > first else rows
        for r in rows:
            if len(r) < 8:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        for r in timed:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples in the upper half of the observed power range
        thr = (max(pw) + min(pw)) / 2 if pw else 0
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        reasons |= set(extra)
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return (d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0),
                "measured (MEASURED_PEAKS.json: sustained bf16 for the tensor kernels timed inside the step, copy bandwidth for HBM)")
    return 6650.0, 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the committed ncu capture,
    valid only for the kernel sources it was taken on (profiles/r02_traffic.json carries their hash); else None."""
    p = ROOT / "profiles" / "r02_traffic.json"
    try:
        d = json.loads(p.read_text())
        import __graft_entry__ as g
        if d.get("csrc_hash") == g.source_hash():
            return d
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------ CPU legs
def _reference_post():
    """The reference's OWN numpy post functions when /root/reference exists (build container), else the pinned
    oracle restatement (GPU box).  Returns (stitch, edit_array, kind)."""
    from oracle import post_oracle as po
    if Path("/root/reference/backend/shared/tts_pipeline.py").exists():
        try:
            from oracle.make_golden import import_reference
            tp, _ae = import_reference()

            def edit(x, sr):     # body of apply_minimal_edit on an array (audio_edit.py:44-69), reference functions inside
                s, e = tp._find_active_range(x, threshold=0.002, min_silence_frames=int(sr * (20 / 1000.0)))
                y = x[s:e] if 0 <= s < e <= len(x) else x
                peak = float(np.max(np.abs(y))) if y.size else 0.0
                if peak > 0.0:
                    y = y * (10 ** (-1.0 / 20.0) / peak)
                return np.clip(y, -1.0, 1.0)
            return (lambda chunks: tp._apply_inter_chunk_gap(chunks, sr=SR, gap_ms=GAP_MS)), edit, "reference numpy functions"
        except Exception:
            pass
    return (lambda chunks: po.apply_inter_chunk_gap(chunks, sr=SR, gap_ms=GAP_MS)), \
        (lambda x, sr: po.apply_minimal_edit_array(x, sr, trim_enabled=True, normalize_enabled=True, target_dbfs=-1.0)[0]), \
        "oracle/post_oracle.py (pinned restatement)"


def cpu_reference_sample(n_chunks, T, threads):
    """The reference's CPU implementation of the path on a bounded sample: torch fp32 restatement of upstream HiFT
    (oracle/hift_oracle.py; chatterbox-tts is not installable offline), chunk by chunk as run_tts_pipeline does, then
    the job post in the reference's order.  Returns (audio_seconds, wall_seconds, post kind)."""
    import torch
    from oracle import hift_oracle as H
    from oracle import post_oracle as po
    torch.set_num_threads(threads)
    stitch, edit, kind = _reference_post()
    W = H.fold_weight_norm(H.make_state_dict(0, "init"))
    mels = [H.synth_mel(T, 2, b) for b in range(n_chunks)]
    pn = [H.synth_noise(T, 2, b) for b in range(n_chunks)]
    t0 = time.perf_counter()
    chunks = [H.hift_inference(mels[b], W, f0=None, phase_vec=pn[b][0], noise=pn[b][1]).numpy() for b in range(n_chunks)]
    raw = po.pcm16_encode(stitch(chunks) if n_chunks > 1 else chunks[0])        # sf.write default subtype (tts_pipeline.py:409)
    out = po.pcm16_encode(edit(po.pcm16_decode(raw), SR))
    dt = time.perf_counter() - t0
    assert out.size > 0
    return n_chunks * T * SPF / SR, dt, kind


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_chunks = 2
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(1, 100, threads)
    times, audio, kind = [], 0.0, ""
    for _ in range(args.steps):
        audio, dt, kind = cpu_reference_sample(n_chunks, T_FRAMES, threads)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = audio / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Chatterbox HiFT batch of {CHUNKS_PER_RANK} chunks x 10 s mel + reference-order job post per GPU "
                               f"(CPU arm: bounded sample of {n_chunks} chunks x 10 s per step)"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "cpu": cpu_model(),
                         "sample": f"{n_chunks} chunks x T={T_FRAMES} per step: torch fp32 HiFT restatement on {threads} threads + {kind}"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_torch_eager(args, rank, world):
    """Secondary baseline: the stock PyTorch-eager path on the SAME B200 - the restated HiFT module
    on cuda (cuDNN convolutions, cuFFT STFT/iSTFT) in fp32 and under bf16 autocast, batch 1 like the reference, plus
    torch ops for the post.  None of this repo's kernels run here."""
    if rank != 0:
        return
    import torch
    from oracle import hift_oracle as H
    dev = torch.device("cuda", 0)
    W = {k: v.to(dev) for k, v in H.fold_weight_norm(H.make_state_dict(0, "init")).items()}
    n_chunks, T = 4, T_FRAMES
    mels = [H.synth_mel(T, 2, b).to(dev) for b in range(n_chunks)]
    pn = [tuple(t.to(dev) for t in H.synth_noise(T, 2, b)) for b in range(n_chunks)]
    out = {}
    for name, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        def step():
            ws = []
            for b in range(n_chunks):
                if ctx is None:
                    ws.append(H.hift_inference(mels[b], W, f0=None, phase_vec=pn[b][0], noise=pn[b][1]))
                else:
                    with ctx:
                        ws.append(H.hift_inference(mels[b], W, f0=None, phase_vec=pn[b][0], noise=pn[b][1]).float())
            x = torch.cat(ws)
            peak = x.abs().max()
            return (x * (10 ** (-1 / 20) / peak)).clamp_(-1, 1)
        for _ in range(max(args.warmup, 2)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[name] = {"value": n_chunks * T * SPF / SR / (ms * 1e-3), "ms_per_step": ms}
    # the best the stock path can do: one batched call (equal-length chunks stack; the reference itself never batches)
    Bn = 16
    melB = torch.stack([H.synth_mel(T, 3, b) for b in range(Bn)]).to(dev)
    pvB = torch.stack([H.synth_noise(T, 3, b)[0] for b in range(Bn)]).to(dev)
    nzB = torch.randn(Bn, 9, T * SPF, device=dev)
    tf = H.trim_fade(torch.float32).to(dev)
    for name, ctx in (("fp32_batch16", None), ("bf16_autocast_batch16", torch.autocast("cuda", dtype=torch.bfloat16))):
        def stepB():
            import contextlib
            with torch.inference_mode(), (ctx or contextlib.nullcontext()):
                f0 = H.f0_predictor(melB, W)
                s_ = H.sine_source(f0.float(), W, pvB, nzB)
                y = H.decode(melB, s_, W).float()
                y[:, : tf.numel()] *= tf
                x = y.reshape(-1)
                return (x * (10 ** (-1 / 20) / x.abs().max())).clamp_(-1, 1)
        for _ in range(max(args.warmup, 2)):
            stepB()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            stepB()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[name] = {"value": Bn * T * SPF / SR / (ms * 1e-3), "ms_per_step": ms}
    best = max(out.values(), key=lambda d: d["value"])
    print(json.dumps({"impl": "torch_eager", "metric": METRIC, "value": best["value"], "unit": "audio-s/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["ms_per_step"], "higher_is_better": True,
                      "data": "synthetic", "dtype": "f32 / bf16 autocast",
                      "config": {"workload": f"stock PyTorch eager (cuDNN/cuFFT) HiFT restatement on cuda:0, 10 s chunks: {n_chunks} chunks at batch 1 per "
                                             f"call like the reference, and one batched call of {Bn} chunks (value = the best variant)"},
                      "variants": out, "gpu_launches": 0}), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def post_sweep(torch, n_samples=268435456, steps=5, warmup=2):
    """BASELINE configs[4]: trim + snap + fades + peak normalise + gap concat over 1 GiB of fp32 audio in 10 s segments
    (per-segment semantics; 12 B per input sample: SURVEY 8(d))."""
    from tools.post_sweep import make_input
    from vocalie_tts_b200 import post
    x, seg_off = make_input(n_samples)
    n_seg = len(seg_off) - 1
    prm = post.make_params(trim=1, min_silence_frames=480, snap_radius=240, fade_in_frames=240, fade_out_frames=240,
                           normalize=1, target_peak=float(10 ** (-1 / 20)), stitch=1, gap_frames=6000, concat=1)
    out = torch.empty(n_samples + n_seg * 6000, dtype=torch.float32, device="cuda")
    for _ in range(warmup):
        post.post_process_device(x, seg_off, prm, out=out, read_back=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        post.post_process_device(x, seg_off, prm, out=out, read_back=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del x, out
    torch.cuda.empty_cache()
    return n_samples, n_seg, ms


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from vocalie_tts_b200.hift import HiFTVocoder, random_state_dict, algorithmic_flops_per_frame
    from vocalie_tts_b200.pipeline import VocoderPipeline
    from vocalie_tts_b200 import distributed as D
    from vocalie_tts_b200.backend import shard_chunks

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    voc = HiFTVocoder(random_state_dict(0), operand=args.operand)
    voc.set_profiling(True)
    pipe = VocoderPipeline(voc, chunk_gap_ms=GAP_MS, out_pcm16=True)       # reference order, PCM_16 out (the WAV the job writes)
    cfg4 = args.workload == "cfg4"
    if cfg4:
        T_all = cfg4_lengths()
        shards = shard_chunks(T_all.tolist(), world)
        max_frames = CFG4_MAX_FRAMES
    else:
        T_all = np.full(args.chunks * world, args.frames, dtype=np.int64)
        shards = D.contiguous_shards(args.chunks * world, world)
        max_frames = None
    Ts = T_all[np.asarray(shards[rank], dtype=np.int64)].astype(np.int32)
    frames_rank = int(Ts.astype(np.int64).sum())
    audio_s_job = float(T_all.sum()) * SPF / SR
    mel_host = pipe.pinned_input(frames_rank)
    mel_host.copy_(synth_mel_host(frames_rank, 1000 + rank))
    mel_dev = mel_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    job = D.ShardedJob(pipe, T_all, shards) if world > 1 else None
    n_raw = D.final_length(T_all * SPF, GAP_MS * SR // 1000)
    finals = [torch.empty(n_raw + 8, dtype=torch.int16, device=dev) for _ in range(2)] if (world > 1 and rank == 0) else [None, None]
    # e2e at N > 1: two host buffers shared by the ranks of the node (/dev/shm, page-locked in every process): every rank
    # copies its own part of the file over its own PCIe link, nothing crosses NVLink and rank 0 reads nothing back
    host_files, host_mode = None, "none"
    if world > 1:
        ok = 1
        try:
            host_files = [D.SharedHostBuffer(f"vocalie_b200_bench_{os.environ.get('MASTER_PORT', '0')}_{i}", n_raw + 8, torch.int16)
                          for i in range(2)]
        except Exception as exc:  # noqa: BLE001  (no /dev/shm or page-locking refused: fall back to the gather + rank-0 read-back)
            print(f"[bench] rank {rank}: shared host buffer unavailable ({exc}); e2e uses the NCCL gather", file=sys.stderr, flush=True)
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            host_files = None
        host_mode = "shared" if host_files else "gather"
    gather_host = [torch.empty(n_raw + 8, dtype=torch.int16).pin_memory() for _ in range(2)] if (host_mode == "gather" and rank == 0) else None
    gather_copied = [torch.cuda.Event(), torch.cuda.Event()]
    gather_stream = torch.cuda.Stream(device=dev) if host_mode == "gather" else None

    def step_device(seed, slot=0):
        if world == 1:
            r = pipe.run_device(mel_dev, Ts, seed=seed, read_back=False, max_frames=max_frames)
            return r, None
        r = job.run_device(mel_dev, seed=seed, out=finals[slot], max_frames=max_frames)
        return r, r.total_samples

    e2e_jobs = [0]

    def e2e_loop(steps, seed0):
        """Host mels in, host PCM_16 audio out, `steps` jobs back to back.  One GPU: the public VocoderPipeline.submit()
        / collect().  Several GPUs: every rank uploads its mels, the sharded job runs, and every rank copies ITS pieces of
        the finished PCM_16 file device -> host into a buffer shared by the ranks (asynchronously, on a side stream, while
        the next job computes; two buffers).  `d2h_bytes_per_step` is the whole file (the sum over ranks)."""
        if world == 1:
            pending = None
            for i in range(steps):
                t = pipe.submit(mel_host, Ts, seed=seed0 + i) if max_frames is None else None
                if t is None:           # bucketed long job: synchronous host API
                    pipe.run(mel_host, Ts, seed=seed0 + i, max_frames=max_frames)
                    continue
                if pending is not None:
                    pipe.collect(pending)
                pending = t
            if pending is not None:
                pipe.collect(pending)
                return pipe.last_d2h_bytes
            return n_raw * 2
        if host_mode == "gather":
            # fallback: device assembly on rank 0 (grouped send / recv), rank 0 reads the file back on a second stream
            cur = torch.cuda.current_stream()
            total = 0
            for i in range(steps):
                slot = i & 1
                if rank == 0:
                    cur.wait_event(gather_copied[slot])
                mel_dev.copy_(mel_host, non_blocking=True)
                r = job.run_device(mel_dev, seed=seed0 + i, out=finals[slot], max_frames=max_frames)
                total = r.total_samples
                if rank == 0:
                    ready = torch.cuda.Event()
                    ready.record(cur)
                    with torch.cuda.stream(gather_stream):
                        gather_stream.wait_event(ready)
                        gather_host[slot][:total].copy_(finals[slot][:total], non_blocking=True)
                        gather_copied[slot].record(gather_stream)
            if gather_stream is not None:
                gather_stream.synchronize()
            cur.synchronize()
            return total * 2
        base = e2e_jobs[0]
        for i in range(steps):
            mel_dev.copy_(mel_host, non_blocking=True)
            r = job.run_device(mel_dev, seed=seed0 + i, max_frames=max_frames, host_out=host_files[(base + i) & 1].tensor)
            if i > 0:
                # job i-1: this rank's copies have landed long ago (they ran under job i's... predecessor's compute); tell
                # the others, and let rank 0 - the consumer of the file - see it complete before its buffer is reused
                job.wait_host(previous=True)
                host_files[(base + i - 1) & 1].publish(base + i - 1)
                if rank == 0:
                    host_files[(base + i - 1) & 1].wait_complete(base + i - 1)
        job.wait_host()
        host_files[(base + steps - 1) & 1].publish(base + steps - 1)
        host_files[(base + steps - 1) & 1].wait_complete(base + steps - 1)
        e2e_jobs[0] = base + steps
        torch.cuda.current_stream().synchronize()
        return r.total_samples * 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the sampler runs from the warm-up on: throttle reasons are collected over warm-up + timed region (the power-cap flag of
    # a ~140 ms timed region is not always set while one of its samples is taken), clocks and power over the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler.mark()
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
        if not cfg4:
            p = voc.read_profile()            # waits for this step's forward; the post kernels are still timed by ev
            for k in prof:
                prof[k] += p[k]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_local = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([ms_local], device=dev)
    per_rank = None
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [float(x.item()) for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = audio_s_job / (ms * 1e-3)

    # ---- end to end through the public host API, every step
    e2e_loop(2 if not cfg4 else 1, 7)
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
    e2e_value = audio_s_job / (float(te.item()) * 1e-3)

    if host_files:
        for hf in host_files:
            hf.close()
    if rank != 0:
        return
    hbm, tf_sus, tf_burst, how = peaks()
    if cfg4:
        mean = statistics.mean(per_rank) if per_rank else ms
        loads = [int(T_all[np.asarray(s, dtype=np.int64)].sum()) for s in shards]
        line = {
            "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.operand] + " operands, f32 accumulate", "data": "synthetic",
            "config": {"workload": f"long-form narration: {CFG4_CHUNKS} chunks of T = 50*U{{1..20}} frames (seed 1004, {audio_s_job:.0f} audio-s), "
                                   f"LPT-sharded by mel length, buckets of <= {CFG4_MAX_FRAMES} frames per vocoder call; post: {pipe.describe()}",
                       "chunks": CFG4_CHUNKS, "frames_per_rank": loads, "l2": "flushed between timed steps (256 MB write)"},
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(sum(loads) * 80 * 4), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clocks,
            "balance": {"per_rank_ms": per_rank or [ms], "max_over_mean": (max(per_rank) / mean) if per_rank else 1.0,
                        "frames_max_over_mean": max(loads) / (sum(loads) / len(loads))},
        }
        print(json.dumps(line), flush=True)
        return
    rb_ms = prof["resblock_ms"] / args.steps
    rb_tflops = prof["resblock_flops"] / args.steps / (rb_ms * 1e-3) / 1e12 if rb_ms > 0 else 0.0
    fwd_ms = prof["total_ms"] / args.steps
    tr = measured_traffic()
    n_chunks, T = args.chunks, args.frames
    line = {
        "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.operand] + " operands, f32 accumulate",
        "data": "synthetic",
        "config": {"workload": f"Chatterbox HiFT batch of {n_chunks} chunks x {T * SPF / SR:g} s mel per GPU + job post: {pipe.describe()}",
                   "chunks_per_gpu": n_chunks, "mel_frames": T, "chunk_gap_ms": GAP_MS, "weights": "random-init (upstream init)",
                   "f0": "predicted (ConvRNNF0Predictor)", "noise": "in-kernel Philox", "output": "PCM_16",
                   "l2": "flushed between timed steps (256 MB write); per-step activations >> L2",
                   "exchange": "int64[3] all-reduce (file trim range + peak), then grouped ncclSend/ncclRecv of every rank's part "
                               "straight into its place on rank 0 (value); e2e: every rank copies its part device->host into "
                               "a host buffer shared by the ranks" + ("" if host_mode == "shared" else " [unavailable here: gather + rank-0 read-back]") if world > 1 else "none"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(mel_host.numel() * 4 * world),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": f"ResBlock convolutions ({prof['resblock_launches'] // args.steps} launches per forward: fused pairs at C=64/128, single convs at C=256)",
                     "achieved": rb_tflops, "peak": tf_sus, "unit": "TFLOP/s", "frac": rb_tflops / tf_sus,
                     "frac_of_burst_peak": rb_tflops / tf_burst,
                     "traffic": tr["dram_bytes_per_launch"] if tr else None,
                     "traffic_source": (tr["source"] + ": " + tr["kernel"]) if tr else "no ncu capture of this build (profiles/r02_traffic.json absent or stale)",
                     "peak_source": how, "kernel_ms_per_step": rb_ms, "forward_ms_per_step": fwd_ms,
                     "kernel_share_of_step": rb_ms / ms if ms > 0 else None,
                     "path_tflops": algorithmic_flops_per_frame() * n_chunks * T * world / (ms * 1e-3) / 1e12},
    }
    if per_rank:
        line["per_rank_ms"] = per_rank
    if world == 1 and not args.no_extras:
        # configs[4]: post-only sweep, the HBM-bound half of the path
        ns, nseg, pms = post_sweep(torch)
        gbs = 12.0 * ns / (pms * 1e-3) / 1e9
        line["roofline_post"] = {"bound": "hbm", "kernel": "k_scan + k_fix + k_peak + k_plan + k_write (trim, snap, fades, peak, gain, 250 ms gap concat)",
                                 "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                                 "workload": f"configs[4]: {ns} fp32 samples (1 GiB) in {nseg} segments of 10 s, 12 B per input sample",
                                 "ms": pms, "audio_s_per_s": ns / SR / (pms * 1e-3)}
        # configs[0]: one ~5 s chunk, host mel in -> host PCM_16 out, p50 of 20
        p1 = VocoderPipeline(voc, chunk_gap_ms=GAP_MS, out_pcm16=True)
        m1 = synth_mel_host(250, 1001).numpy()
        T1 = np.array([250], np.int32)
        lat = []
        for i in range(24):
            w0 = time.perf_counter()
            p1.run(m1, T1, seed=i)
            lat.append((time.perf_counter() - w0) * 1e3)
        lat = sorted(lat[4:])
        line["latency_cfg1"] = {"workload": "configs[0]: B=1, T=250 (5 s), host mel -> host PCM_16, trim + -1 dBFS normalise",
                                "p50_ms": lat[len(lat) // 2], "p90_ms": lat[int(len(lat) * 0.9)], "real_time_factor": 5000.0 / lat[len(lat) // 2]}
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        cpu_reference_sample(1, 250, threads)      # warm-up (configs[0] shape)
        a, dt, kind = cpu_reference_sample(4, T_FRAMES, threads)
        line["cpu_baseline"] = {"value": a / dt, "unit": "audio-s/s", "cores": threads, "kind": "port", "cpu": cpu_model(),
                                "sample": f"4 chunks x T={T_FRAMES} (40 s audio): torch fp32 HiFT restatement on {threads} threads + {kind}"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_eager"])
    ap.add_argument("--operand", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"])
    ap.add_argument("--chunks", type=int, default=CHUNKS_PER_RANK)
    ap.add_argument("--frames", type=int, default=T_FRAMES)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg5 sweep and the cfg1 latency (N = 1 line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.impl == "torch_eager":
        run_torch_eager(args, rank, world)
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
