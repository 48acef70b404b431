"""Host side of the B200 HiFT vocoder: weight folding/packing, ragged batching, the ctypes call.

Mirrors the upstream call the reference reaches through ``tts.generate`` (reference
tts_backends/chatterbox_impl.py:189): ``HiFTGenerator.inference(speech_feat=mel)`` followed by
the ``S3Token2Wav`` ``trim_fade`` tail - mel ``[80, T]`` in, waveform ``[480*T]`` at 24 kHz out -
batched over independent chunks.  All arithmetic runs in ``csrc/`` through the C ABI
(``include/vocalie_b200.h``); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check
from .errors import BackendUnavailableError

S3GEN_SR = 24000            # upstream const.py
N_MEL = 80
SAMPLES_PER_FRAME = 480     # 8 * 5 * 3 upsampling x hop 4
N_HARMONICS = 9

OPERANDS = {"fp16": _lib.VT_OPERAND_FP16, "bf16": _lib.VT_OPERAND_BF16, "fp32": _lib.VT_OPERAND_FP32}

# Constructor arguments of upstream HiFTGenerator per consumer (the module class is the same: Chatterbox's hifigan.py is
# CosyVoice's cosyvoice/hifigan/generator.py).  "chatterbox_s3gen": S3Token2Wav.__init__.  "cosyvoice_300m":
# the `hift:` block of CosyVoice-300M's cosyvoice.yaml - the engine behind tts_backends/cosyvoice_runner.py:75-131 at its
# default 22 050 Hz (:84,131).
GENERATOR_CONFIGS = {
    "chatterbox_s3gen": dict(sampling_rate=24000, upsample_rates=(8, 5, 3), upsample_kernel_sizes=(16, 11, 7),
                             source_resblock_kernel_sizes=(7, 7, 11), trim_fade=True),
    "cosyvoice_300m": dict(sampling_rate=22050, upsample_rates=(8, 8), upsample_kernel_sizes=(16, 16),
                           source_resblock_kernel_sizes=(7, 11), trim_fade=False),
}


def resolve_config(config) -> dict:
    if config is None:
        config = "chatterbox_s3gen"
    if isinstance(config, str):
        if config not in GENERATOR_CONFIGS:
            raise ValueError(f"unknown generator config {config!r} (known: {sorted(GENERATOR_CONFIGS)})")
        config = GENERATOR_CONFIGS[config]
    cfg = dict(GENERATOR_CONFIGS["chatterbox_s3gen"], **{k: config[k] for k in config})
    n = len(cfg["upsample_rates"])
    if n not in (2, 3) or len(cfg["upsample_kernel_sizes"]) != n or len(cfg["source_resblock_kernel_sizes"]) != n:
        raise ValueError("generator config: 2 or 3 upsampling stages with matching kernel-size lists are supported")
    return cfg


def source_down_shapes(upsample_rates):
    """(k, stride, pad) of source_downs[i] - upstream HiFTGenerator.__init__: stride = product of the later rates;
    1 -> Conv1d(k=1), else Conv1d(k=2u, stride=u, padding=u//2)."""
    out = []
    for i in range(len(upsample_rates)):
        u = int(np.prod(upsample_rates[i + 1:])) if i + 1 < len(upsample_rates) else 1
        out.append((1, 1, 0) if u == 1 else (2 * u, u, u // 2))
    return out


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise BackendUnavailableError("no CUDA device: the B200 HiFT path has no CPU fallback")
    return torch


def fold_weight_norm(state_dict: Dict[str, "object"]) -> Dict[str, np.ndarray]:
    """Fold ``torch.nn.utils.parametrizations.weight_norm`` (``w = g * v / ||v||`` over all dims
    but 0) and return float32 numpy arrays keyed by plain upstream names.  Entries
    that are already plain ``.weight`` pass through; legacy ``weight_g``/``weight_v`` keys are
    folded too."""
    def to_np(v):
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        return np.asarray(v)

    out: Dict[str, np.ndarray] = {}
    pairs = ((".parametrizations.weight.original0", ".parametrizations.weight.original1"), (".weight_g", ".weight_v"))
    for k, v in state_dict.items():
        done = False
        for gs, vs in pairs:
            if k.endswith(vs):
                base = k[: -len(vs)]
                g = to_np(state_dict[base + gs]).astype(np.float64)
                vv = to_np(v).astype(np.float64)
                norm = np.sqrt((vv.reshape(vv.shape[0], -1) ** 2).sum(axis=1)).reshape((-1,) + (1,) * (vv.ndim - 1))
                out[base + ".weight"] = np.ascontiguousarray((g.reshape(norm.shape) * vv / norm).astype(np.float32))
                done = True
            elif k.endswith(gs):
                done = True
        if not done:
            out[k] = np.ascontiguousarray(to_np(v).astype(np.float32))
    return out


def algorithmic_flops_per_frame(include_f0: bool = True, config=None) -> float:
    """2*MAC of every conv of the path per mel frame (612.45 MFLOP with the F0 predictor for Chatterbox)."""
    cfg = resolve_config(config)
    rates, kernels, src_k = cfg["upsample_rates"], cfg["upsample_kernel_sizes"], cfg["source_resblock_kernel_sizes"]
    f = 0.0
    if include_f0:
        f += 2 * (80 * 512 * 3 + 4 * 512 * 512 * 3 + 512)
    f += 2 * 80 * 512 * 7
    steps = 1                                   # input steps per frame of the stage's transposed conv
    sds = source_down_shapes(rates)
    for i, (u, k) in enumerate(zip(rates, kernels)):
        cin, c = 512 >> i, 512 >> (i + 1)
        f += 2 * cin * c * k * steps
        steps *= u
        f += 2 * 18 * c * sds[i][0] * steps
        f += 6 * 2 * c * c * src_k[i] * steps
        for rk in (3, 7, 11):
            f += 6 * 2 * c * c * rk * steps
    f += 2 * (512 >> len(rates)) * 18 * 7 * steps
    return f


class HiFTVocoder:
    """A HiFT generator resident on the current CUDA device.

    ``state_dict``: upstream ``HiFTGenerator`` state dict (weight-norm parametrised or folded).
    ``operand``: "fp16" (default; tensor-core operands, fp32 accumulate - meets the 1e-3 / 60 dB
    parity bar), "bf16", or "fp32" (exact CUDA-core path).
    """

    def __init__(self, state_dict, operand: str = "fp16", config=None):
        """``config``: a name of ``GENERATOR_CONFIGS`` or a dict of HiFTGenerator constructor arguments
        (default: Chatterbox S3Gen)."""
        torch = _torch()
        if operand not in OPERANDS:
            raise ValueError(f"operand must be one of {sorted(OPERANDS)}")
        self.operand = operand
        self.config = resolve_config(config)
        self._lib = _lib.load_library()
        sm, major, minor = C.c_int(), C.c_int(), C.c_int()
        check(self._lib.vt_device_check(C.byref(sm), C.byref(major), C.byref(minor)), "vt_device_check")
        self.sm_count = sm.value
        folded = fold_weight_norm(state_dict)
        names = [k for k in folded if not k.endswith("num_batches_tracked")]
        arr = (_lib.Tensor * len(names))()
        keep = []
        for i, k in enumerate(names):
            a = folded[k]
            if a.ndim == 0:
                a = a.reshape(1)
            if a.ndim > 4:
                raise ValueError(f"{k}: rank {a.ndim} tensors are not part of HiFT")
            keep.append((k.encode(), a))
            arr[i].name = keep[-1][0]
            arr[i].data = a.ctypes.data_as(C.c_void_p)
            arr[i].ndim = a.ndim
            for d in range(a.ndim):
                arr[i].shape[d] = a.shape[d]
        h = C.c_void_p()
        cc = _lib.HiftConfig()
        cc.sampling_rate = int(self.config["sampling_rate"])
        cc.n_upsamples = len(self.config["upsample_rates"])
        for i in range(cc.n_upsamples):
            cc.upsample_rates[i] = int(self.config["upsample_rates"][i])
            cc.upsample_kernel_sizes[i] = int(self.config["upsample_kernel_sizes"][i])
            cc.source_resblock_kernel_sizes[i] = int(self.config["source_resblock_kernel_sizes"][i])
        cc.trim_fade = 1 if self.config["trim_fade"] else 0
        check(self._lib.vt_hift_create_ex(arr, len(names), OPERANDS[operand], C.byref(cc), C.byref(h)), "vt_hift_create_ex")
        self._h = h
        self.samples_per_frame = int(self._lib.vt_hift_samples_per_frame(h))
        self.sr = int(self._lib.vt_hift_sampling_rate(h))
        self._ws = None
        self._lock = threading.Lock()   # reference jobs run on up to 2 threads (backend/config.py:11)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.last_launches = 0
        self._last_call_launches = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vt_hift_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ batching helpers
    @staticmethod
    def pack_mels(mels: Sequence, device=None):
        """List of upstream-layout mels ``[80, T_b]`` -> (frame-major float32 ``[sum_T, 80]``, int32 T)."""
        torch = _torch()
        T = np.array([int(m.shape[-1]) for m in mels], dtype=np.int32)
        rows = [m.reshape(N_MEL, -1).to(device or "cuda", torch.float32).t() for m in mels]
        packed = torch.cat(rows, dim=0).contiguous() if rows else torch.zeros((0, N_MEL), device="cuda")
        return packed, T

    def workspace_bytes(self, B: int, total_T: int, T_max: int) -> int:
        n = int(self._lib.vt_hift_workspace_bytes(self._h, B, total_T, T_max))
        if n < 0:
            raise BackendUnavailableError("vt_hift_workspace_bytes failed")
        return n

    def _workspace(self, torch, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    # ------------------------------------------------------------------ forward
    def forward_packed(self, mel, T, *, f0=None, phase_vec=None, noise=None, seed: int = 0, out=None):
        """``mel``: float32 CUDA ``[sum_T, 80]`` frame-major; ``T``: int32 host array of frames per
        sequence.  Optional ``f0`` ``[sum_T]`` (Hz), ``phase_vec`` ``[B, 9]``, ``noise`` packed
        ``[9, 480*T_b]`` blocks (see the header).  Returns the packed waveform ``[480*sum_T]``
        (sequence b at ``480 * sum(T[:b])``)."""
        torch = _torch()
        T = np.ascontiguousarray(T, dtype=np.int32)
        B = int(T.size)
        total_T = int(T.sum())
        if mel.dtype != torch.float32 or not mel.is_cuda or mel.numel() != total_T * N_MEL:
            raise ValueError("mel must be a float32 CUDA tensor of shape [sum(T), 80]")
        mel = mel.contiguous()
        for name, t, n in (("f0", f0, total_T), ("phase_vec", phase_vec, B * N_HARMONICS),
                           ("noise", noise, total_T * self.samples_per_frame * N_HARMONICS)):
            if t is not None and (t.dtype != torch.float32 or not t.is_cuda or t.numel() != n or not t.is_contiguous()):
                raise ValueError(f"{name} must be a contiguous float32 CUDA tensor with {n} elements")
        if out is None:
            out = torch.empty(total_T * self.samples_per_frame, dtype=torch.float32, device=self.device)
        elif out.numel() < total_T * self.samples_per_frame or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError("out must be a float32 CUDA tensor with 480*sum(T) elements")
        # launch on the vocoder's device and on the calling thread's current stream FOR THAT DEVICE (reference job threads
        # default to device 0 whatever device the vocoder lives on)
        with self._lock, torch.cuda.device(self.device):
            ws = self._workspace(torch, self.workspace_bytes(B, total_T, int(T.max()) if B else 0))
            ptr = lambda t: 0 if t is None else int(t.data_ptr())
            rc = self._lib.vt_hift_forward(self._h, ptr(mel), T.ctypes.data_as(C.POINTER(C.c_int32)), B, ptr(f0),
                                           ptr(phase_vec), ptr(noise), int(seed) & (2 ** 64 - 1), ptr(out), ptr(ws),
                                           ws.numel(), int(torch.cuda.current_stream().cuda_stream))
            check(rc, "vt_hift_forward")
            self.last_launches = self._last_call_launches = int(self._lib.vt_last_launch_count())
        return out

    def forward_bucketed(self, mel, T, *, f0=None, phase_vec=None, noise=None, seed: int = 0, out=None,
                         max_frames: Optional[int] = None):
        """``forward_packed`` over consecutive buckets of sequences holding at most ``max_frames`` mel frames each
        (a single longer sequence gets a bucket of its own): the workspace is sized by the bucket, not by the job, so
        a 2 048-chunk narration (BASELINE configs[3]) runs in bounded memory.  The ragged batch is packed, not padded,
        so bucketing costs no arithmetic; results are identical to one call (sequences are independent; the in-kernel
        generator is keyed on the seed and the sequence's index IN ITS BUCKET, so pass explicit ``phase_vec`` /
        ``noise`` when bit-equality with an unbucketed call matters)."""
        torch = _torch()
        T = np.ascontiguousarray(T, dtype=np.int32)
        total_T = int(T.astype(np.int64).sum())
        if max_frames is None or total_T <= max_frames or T.size <= 1:
            return self.forward_packed(mel, T, f0=f0, phase_vec=phase_vec, noise=noise, seed=seed, out=out)
        if out is None:
            out = torch.empty(total_T * self.samples_per_frame, dtype=torch.float32, device=self.device)
        off = np.concatenate([[0], np.cumsum(T.astype(np.int64))])
        b0 = 0
        launches = 0
        while b0 < T.size:
            b1 = b0 + 1
            while b1 < T.size and off[b1 + 1] - off[b0] <= max_frames:
                b1 += 1
            f_lo, f_hi = int(off[b0]), int(off[b1])
            self.forward_packed(mel[f_lo:f_hi], T[b0:b1], f0=None if f0 is None else f0[f_lo:f_hi],
                                phase_vec=None if phase_vec is None else phase_vec.reshape(-1, N_HARMONICS)[b0:b1].contiguous(),
                                noise=None if noise is None else noise.reshape(-1)[f_lo * self.samples_per_frame * N_HARMONICS:
                                                                                  f_hi * self.samples_per_frame * N_HARMONICS],
                                seed=seed + b0, out=out[f_lo * self.samples_per_frame:f_hi * self.samples_per_frame])
            launches += self.last_launches
            b0 = b1
        self.last_launches = launches          # all buckets (the library's own counter restarts with every call)
        return out

    def inference(self, mels: Sequence, *, f0: Optional[Sequence] = None, phase_vec=None, noise: Optional[Sequence] = None,
                  seed: int = 0) -> List:
        """Batched ``HiFTGenerator.inference``: list of mels ``[80, T_b]`` -> list of waveforms ``[480*T_b]``."""
        torch = _torch()
        mel, T = self.pack_mels(mels, self.device)
        f0p = None if f0 is None else torch.cat([x.reshape(-1).to(self.device, torch.float32) for x in f0]).contiguous()
        pv = None if phase_vec is None else torch.stack([p.reshape(-1) for p in phase_vec]).to(self.device, torch.float32).contiguous()
        nz = None if noise is None else torch.cat([n.reshape(-1).to(self.device, torch.float32) for n in noise]).contiguous()
        wav = self.forward_packed(mel, T, f0=f0p, phase_vec=pv, noise=nz, seed=seed)
        off = np.concatenate([[0], np.cumsum(T.astype(np.int64) * self.samples_per_frame)])
        return [wav[off[i]:off[i + 1]] for i in range(len(T))]

    def set_profiling(self, enable: bool = True):
        check(self._lib.vt_hift_set_profiling(self._h, 1 if enable else 0), "vt_hift_set_profiling")

    def read_profile(self) -> dict:
        """Device timing of the last forward: {'total_ms', 'resblock_ms', 'resblock_flops', 'resblock_launches'}."""
        t, r, f, n = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        check(self._lib.vt_hift_read_profile(self._h, C.byref(t), C.byref(r), C.byref(f), C.byref(n)), "vt_hift_read_profile")
        return {"total_ms": t.value, "resblock_ms": r.value, "resblock_flops": f.value, "resblock_launches": n.value}

    def read_timeline(self) -> Dict[str, float]:
        """Device milliseconds per section of the last forward (profiling must be on)."""
        buf = C.create_string_buffer(4096)
        check(self._lib.vt_hift_read_timeline(self._h, buf, len(buf)), "vt_hift_read_timeline")
        out: Dict[str, float] = {}
        for line in buf.value.decode().splitlines():
            k, v = line.split("=")
            out[k] = float(v)
        return out

    def read_tap(self, name: str, seq: int, channels: int = 1):
        """Intermediate of the last forward as float32 ``[rows, channels]`` (parity tests)."""
        torch = _torch()
        st = int(torch.cuda.current_stream().cuda_stream)
        n = int(self._lib.vt_hift_read_tap(self._h, name.encode(), int(seq), 0, 0, 0, st))
        if n < 0:
            check(n, f"vt_hift_read_tap({name})")
        out = torch.empty(max(n, 1), dtype=torch.float32, device=self.device)
        n2 = int(self._lib.vt_hift_read_tap(self._h, name.encode(), int(seq), int(out.data_ptr()), out.numel(), 0, st))
        if n2 < 0:
            check(n2, f"vt_hift_read_tap({name})")
        return out[:n].view(-1, channels)


def random_state_dict(seed: int = 0, config=None) -> Dict[str, np.ndarray]:
    """Random-init HiFT weights in upstream state-dict naming (benchmarks and smoke tests - there
    is no network to fetch ``ResembleAI/chatterbox``'s ``s3gen`` checkpoint).  Follows upstream
    initialisation (SURVEY 8(d) W_init): ``normal(0, 0.01)`` on the weight-normed convs of ups /
    resblocks / source_resblocks / conv_post, ``U(+-1/sqrt(fan_in))`` elsewhere, Snake alpha = 1,
    weight-norm gain ``g = ||v||``."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, np.ndarray] = {}
    cfg = resolve_config(config)
    rates, up_k, src_k = cfg["upsample_rates"], cfg["upsample_kernel_sizes"], cfg["source_resblock_kernel_sizes"]
    sds = source_down_shapes(rates)

    def conv(name, cout, cin, k, normal, wn=True, transposed=False):
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        fan_in = (cout if transposed else cin) * k
        if normal:
            v = rng.standard_normal(shape).astype(np.float32) * np.float32(0.01)
        else:
            v = rng.uniform(-1, 1, shape).astype(np.float32) / np.float32(np.sqrt(fan_in))
        if wn:
            sd[name + ".parametrizations.weight.original1"] = v
            sd[name + ".parametrizations.weight.original0"] = np.sqrt((v.reshape(v.shape[0], -1) ** 2).sum(1)).reshape(-1, 1, 1).astype(np.float32)
        else:
            sd[name + ".weight"] = v
        sd[name + ".bias"] = (rng.uniform(-1, 1, cout).astype(np.float32) / np.float32(np.sqrt(fan_in)))

    conv("conv_pre", 512, 80, 7, False)
    for i, k in enumerate(up_k):
        cout = 512 >> (i + 1)
        conv(f"ups.{i}", cout, 512 >> i, k, True, transposed=True)
        conv(f"source_downs.{i}", cout, 18, sds[i][0], False, wn=False)
        for j in range(3):
            for c12 in ("convs1", "convs2"):
                conv(f"source_resblocks.{i}.{c12}.{j}", cout, cout, src_k[i], True)
            for a in ("activations1", "activations2"):
                sd[f"source_resblocks.{i}.{a}.{j}.alpha"] = np.ones(cout, np.float32)
        for kk, rk in enumerate((3, 7, 11)):
            r = i * 3 + kk
            for j in range(3):
                for c12 in ("convs1", "convs2"):
                    conv(f"resblocks.{r}.{c12}.{j}", cout, cout, rk, True)
                for a in ("activations1", "activations2"):
                    sd[f"resblocks.{r}.{a}.{j}.alpha"] = np.ones(cout, np.float32)
    conv("conv_post", 18, 512 >> len(rates), 7, True)
    for i in range(5):
        conv(f"f0_predictor.condnet.{2 * i}", 512, 80 if i == 0 else 512, 3, False)
    sd["f0_predictor.classifier.weight"] = rng.uniform(-1, 1, (1, 512)).astype(np.float32) / np.float32(np.sqrt(512))
    sd["f0_predictor.classifier.bias"] = rng.uniform(-1, 1, 1).astype(np.float32) / np.float32(np.sqrt(512))
    sd["m_source.l_linear.weight"] = rng.uniform(-1, 1, (1, 9)).astype(np.float32) / np.float32(3.0)
    sd["m_source.l_linear.bias"] = rng.uniform(-1, 1, 1).astype(np.float32) / np.float32(3.0)
    return sd
