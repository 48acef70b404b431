/*
 * vocalie_b200.h - C ABI of the B200-native HiFT vocoder + post-processing path.
 *
 * Drop-in boundary for the hot path of Bricesodini/Vocalie-TTS named in BASELINE.json:
 * everything below `TTSBackend.synthesize_chunk` (reference tts_backends/base.py:190-217,
 * tts_backends/chatterbox_backend.py:176-192) from mel to finished audio, i.e. upstream
 * chatterbox-tts==0.1.6 `HiFTGenerator.inference` (call site tts_backends/chatterbox_impl.py:189)
 * and the numpy post-processing of backend/shared/tts_pipeline.py:114-274 and
 * backend/shared/audio_edit.py:16-79.
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch types.  Unless a parameter is marked HOST, every
 *    pointer is a DEVICE pointer owned by the caller; nothing is allocated behind the caller's
 *    back except inside handles created by vt_*_create (freed by vt_*_destroy).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All calls are
 *    asynchronous on that stream unless stated otherwise.
 *  - every function returns VT_OK (0) or a negative VT_ERR_* code; vt_last_error() returns a
 *    thread-local message.  The Python shim turns failures into BackendUnavailableError
 *    (reference tts_backends/base.py:220) - the single error type at the boundary.
 *  - audio is mono float32; segments ("chunks" in the reference) of a batch are described by
 *    an offsets array seg_off[n_seg+1] of int64 sample positions into one flat buffer.
 */
#ifndef VOCALIE_B200_H
#define VOCALIE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VT_OK               0
#define VT_ERR_INVALID     -1   /* bad argument */
#define VT_ERR_CUDA        -2   /* CUDA runtime error (message has cudaGetErrorString) */
#define VT_ERR_UNSUPPORTED -3   /* not an sm_100 device / unsupported shape */
#define VT_ERR_NOMEM       -4

#define VT_ABI_VERSION 3

/* ---- library ------------------------------------------------------------------------- */
int         vt_abi_version(void);
const char* vt_last_error(void);
/* Fills sm_count / compute capability of the current device; VT_ERR_UNSUPPORTED if not 10.x. */
int         vt_device_check(int* sm_count, int* cc_major, int* cc_minor);

/* ---- post-processing: reference backend/shared/tts_pipeline.py ----------------------- */

/* Common arguments of the segment-batched calls:
 *   audio       : float32 flat buffer (16-byte aligned) holding all segments
 *   seg_off     : int64[n_seg+1] sample offsets of the segments in `audio` (n_seg <= 65535)
 *   n_samples   : HOST copy of seg_off[n_seg];  max_seg_len: HOST max segment length
 *   workspace   : >= vt_post_workspace_bytes(n_seg, n_samples) bytes, 256-byte aligned
 */
int64_t vt_post_workspace_bytes(int n_seg, int64_t n_samples);

/* _find_active_range (tts_pipeline.py:192-209), batched over segments.
 * ranges[2*i] = start, ranges[2*i+1] = end (relative to the segment).  fp32 compare
 * |x| > threshold; all-silent -> (0, len); start < min_silence -> 0; len-end < min_silence -> len;
 * empty segment -> (0, 0). */
int vt_find_active_range(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples,
                         int64_t max_seg_len, float threshold, int min_silence_frames,
                         int64_t* ranges, void* workspace, int64_t workspace_bytes, void* stream);

/* Raw statistics of every segment in one read pass: first_last[2*i], first_last[2*i+1] = first and last sample
 * with |x| > threshold (segment-relative; -1, -1 when there is none), peak[i] = max|x| over the segment.  This is
 * what the ranks of a sharded job exchange (min / max / max all-reduce of three scalars) to reproduce the whole-file
 * trim and the single peak of apply_minimal_edit (backend/services/tts_service.py:195-207, audio_edit.py:44-66):
 * the min-silence rule of _find_active_range is then applied once, on the global range. */
int vt_post_stats(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples,
                  int64_t max_seg_len, float threshold, int64_t* first_last, float* peak,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* _snap_zero_crossing (tts_pipeline.py:114-137), batched: idx_out[i] = snapped idx_in[i]
 * inside segment i (radius inclusive, ties -> lower index, none -> clamped idx). */
int vt_snap_zero_crossing(const float* audio, const int64_t* seg_off, int n_seg,
                          const int64_t* idx_in, int radius_samples,
                          int64_t* idx_out, void* stream);

/* Segment-batched post-processing plan.  One parameter block covers the three reference bodies:
 *   minimal_post_process  (tts_pipeline.py:212-274): trim=1 snap=240 fades in->out, per-segment peak
 *   apply_minimal_edit    (audio_edit.py:16-79)    : trim=0/1 snap=-1 no fades, peak gain, clip
 *   _apply_inter_chunk_gap(tts_pipeline.py:162-189): stitch=1, edge fades out->in, gap zeros
 */
typedef struct vt_post_params {
  int32_t sr;                 /* sample rate (informational) */
  int32_t trim;               /* 1: _find_active_range trim per segment */
  float   silence_threshold;  /* 0.002f (audio_defaults.py:3) */
  int32_t min_silence_frames; /* int(sr*(int(20)/1000.0)) = 480 */
  int32_t snap_radius;        /* >=0: zero-cross snap of both ends (240); <0: no snap */
  int32_t fade_in_frames;     /* _fade_in length at the head of every output segment (240); 0: none */
  int32_t fade_out_frames;    /* _fade_out length at the tail (240); 0: none */
  int32_t stitch;             /* 1: _apply_inter_chunk_gap semantics: no fade-in on the first
                                 segment, no fade-out on the last, fade-out applied before fade-in */
  int32_t gap_frames;         /* zeros inserted between consecutive segments (stitch only) */
  int32_t normalize;          /* 0: none; 1: per-segment peak; 2: one peak over all segments */
  int32_t clip;               /* 1: clip to [-1, 1] after gain (audio_edit.py:69) */
  double  target_peak;        /* 10**(dBFS/20), computed by the caller in float64 */
  int32_t concat;             /* 1: outputs packed back to back (+gaps) in segment order;
                                 0: segment i is written at out + seg_off[i] (same layout as input) */
  int32_t out_pcm16;          /* 1: out is int16 PCM (lrintf(x*32767)), 0: float32 */
  int32_t stitch_head;        /* stitch only: 1 = segment 0 is the first chunk of the whole job (no fade-in) */
  int32_t stitch_tail;        /* stitch only: 1 = the last segment is the last chunk of the whole job (no
                                 fade-out, no trailing gap).  A rank holding a middle shard of a sharded job
                                 passes 0/0: every segment is faded on both sides and followed by a gap, so
                                 the concatenation of the ranks' outputs equals the single-GPU result. */
} vt_post_params;

/* Per-segment results, 8 doubles per segment (device array `results`, may be NULL):
 *   [0] start  [1] end  (after snap/fallback, relative to the segment)
 *   [2] peak_before (fp32 max|x| of the trimmed+faded segment)  [3] scale (float64)
 *   [4] dst offset (samples)  [5] output length  [6] peak used for the gain  [7] reserved
 * total_out (device int64[1], may be NULL) receives the total number of output samples. */
#define VT_POST_RESULT_STRIDE 8

/* Pass 1: ranges, snap, fade lengths, peaks -> plan kept in `workspace`.
 * range_override (device int64[n_seg][2], nullable) replaces the trim analysis. */
int vt_post_analyze(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples,
                    int64_t max_seg_len, const vt_post_params* params /* HOST */,
                    const int64_t* range_override, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* Pass 2: gains, output offsets, fade*gain*gap write (float32 or PCM_16).  Must follow
 * vt_post_analyze on the same workspace.  peak_override (device float[1], nullable) replaces the
 * analysed peak - the hook for a cross-rank max (whole-file normalisation of a sharded job). */
int vt_post_write(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples,
                  int64_t max_seg_len, const vt_post_params* params /* HOST */,
                  const float* peak_override, void* out, int64_t out_capacity_samples,
                  double* results, int64_t* total_out, void* workspace, int64_t workspace_bytes,
                  void* stream);

/* vt_post_analyze followed by vt_post_write. */
int vt_post_process(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples,
                    int64_t max_seg_len, const vt_post_params* params /* HOST */,
                    void* out, int64_t out_capacity_samples, double* results, int64_t* total_out,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* PCM_16 wire format (tts_backends/chatterbox_runner.py:152 sf.write default subtype,
 * tts_backends/base_runner.py:323 sf.read(dtype="float32")). */
int vt_pcm16_encode(const float* in, int16_t* out, int64_t n, void* stream);
int vt_pcm16_decode(const int16_t* in, float* out, int64_t n, void* stream);

/* The 44-byte RIFF/WAVE header of a mono PCM_16 file, written on the device in front of the samples vt_post_write
 * produced (out_pcm16 = 1), so a finished - possibly trimmed - job leaves the GPU as a file image in one copy.  The
 * sample count is *n_samples_dev (device int64, e.g. the writer's total_out) when non-NULL, else n_samples_host.
 * Byte-identical to what the reference's writers produce for these parameters (tts_backends/chatterbox_runner.py:152,
 * backend/shared/tts_pipeline.py:409, backend/shared/audio_edit.py:70). */
#define VT_WAV_HEADER_BYTES 44
int vt_wav_pcm16_header(void* dst /* device, 44 bytes */, int sample_rate, const int64_t* n_samples_dev,
                        int64_t n_samples_host, void* stream);

/* _resample_audio (backend/shared/tts_pipeline.py:100-111 -> librosa.resample), segment-batched rational-ratio polyphase
 * resampler: segment i of `in` (seg_off_in) -> segment i of `out` (seg_off_out, lengths by librosa's rule
 * int(ceil(n * target_sr / orig_sr)), computed by the caller), up / down = target_sr / orig_sr reduced.
 * table: device float [up][ntaps] phase table (ntaps odd), y[m] = sum_j table[(m*down) % up][j] * x[(m*down)/up + (ntaps-1)/2 - j].
 * The filter itself is the caller's (vocalie-tts_b200/post.py builds a 120 dB Kaiser sinc); sample values are NOT pinned
 * to soxr, which the reference reaches through librosa and which is absent here (oracle/resample_oracle.py). */
int vt_resample(const float* in, const int64_t* seg_off_in, const int64_t* seg_off_out, int n_seg,
                int64_t max_out_len /* HOST */, int up, int down, const float* table, int ntaps, float* out,
                void* stream);

/* RMS helper the reference uses to validate clips (tts_backends/cosyvoice_backend.py:103,
 * tests/test_qwen3_runner.py:58): rms_out[i] = sqrt(mean(float64(x)^2)) over segment i, 0 for an empty
 * segment.  float64 accumulation in a fixed reduction order (deterministic; equal to numpy's pairwise
 * sum to ~1e-15 relative).  workspace >= VT_RMS_PARTIALS * n_seg doubles. */
#define VT_RMS_PARTIALS 64
int vt_rms(const float* audio, const int64_t* seg_off, int n_seg, double* rms_out, void* workspace,
           int64_t workspace_bytes, void* stream);

/* ---- HiFT vocoder: upstream chatterbox/models/s3gen/hifigan.py ------------------------- */

typedef struct vt_hift vt_hift;   /* opaque handle: packed weights + layer plan */

/* Operand precision of the tensor-core convolutions (accumulation is always fp32). */
#define VT_OPERAND_FP16 0   /* fp16 operands on tcgen05 (default: meets the 60 dB parity bar) */
#define VT_OPERAND_BF16 1   /* bf16 operands on tcgen05 */
#define VT_OPERAND_FP32 2   /* exact path: every layer in fp32 on CUDA cores */

/* One entry of the flat weight table handed to vt_hift_create: upstream state-dict name,
 * HOST float32 data with weight-norm already folded (w = g*v/||v||), and shape
 * (conv: [C_out, C_in, k]; ConvTranspose1d: [C_in, C_out, k]; vectors: [n]). */
typedef struct vt_tensor {
  const char*  name;
  const float* data;     /* HOST */
  int32_t      ndim;
  int64_t      shape[4];
} vt_tensor;

int  vt_hift_create(const vt_tensor* tensors /* HOST */, int n_tensors, int operand_dtype,
                    vt_hift** out_handle);

/* The constructor arguments of upstream HiFTGenerator that differ between its users (the module class is one and the
 * same: Chatterbox's hifigan.py is CosyVoice's cosyvoice/hifigan/generator.py).  NULL / vt_hift_create = Chatterbox S3Gen
 * (24 kHz, rates 8/5/3, kernels 16/11/7, source ResBlock kernels 7/7/11, trim_fade tail).  The second consumer behind the
 * reference's plugin interface is CosyVoice (tts_backends/cosyvoice_runner.py:75-131, default rate 22 050 Hz :84,131):
 * CosyVoice-300M = {22050, 2, {8, 8}, {16, 16}, {7, 11}, 0}.  Channels are 512 >> stage, n_fft 16 / hop 4, ResBlock
 * kernels 3/7/11 with dilations 1/3/5, nine harmonics - as in every published instantiation of this generator.
 * Layers whose shape has a tcgen05 instance run on it, the others on the fp32 CUDA-core kernels. */
typedef struct vt_hift_config {
  int32_t sampling_rate;
  int32_t n_upsamples;                       /* 2 or 3 */
  int32_t upsample_rates[4];
  int32_t upsample_kernel_sizes[4];
  int32_t source_resblock_kernel_sizes[4];   /* 3, 7 or 11 each */
  int32_t trim_fade;                         /* 1: S3Token2Wav tail (first sr/50 samples zeroed, next sr/50 faded in) */
} vt_hift_config;
int  vt_hift_create_ex(const vt_tensor* tensors /* HOST */, int n_tensors, int operand_dtype,
                       const vt_hift_config* config /* HOST, nullable */, vt_hift** out_handle);
int  vt_hift_samples_per_frame(const vt_hift* h);   /* prod(upsample_rates) * hop: 480 (Chatterbox), 256 (CosyVoice-300M) */
int  vt_hift_sampling_rate(const vt_hift* h);
void vt_hift_destroy(vt_hift* h);

/* Bytes of device workspace vt_hift_forward needs for a batch of B sequences whose mel lengths
 * sum to total_T with maximum T_max. */
int64_t vt_hift_workspace_bytes(const vt_hift* h, int B, int64_t total_T, int64_t T_max);

/* HiFTGenerator.inference + the S3Token2Wav trim_fade tail, batched over ragged sequences.
 *   mel      : float32 [sum_T, 80] frame-major (row t of sequence b at mel_off[b]+t)
 *   T        : HOST int32[B] frames per sequence;
 *   f0       : float32 [sum_T] Hz or NULL (run the ConvRNNF0Predictor)
 *   phase_vec: float32 [B, 9] or NULL;  noise: float32 [9, L] per sequence packed at 9*wav_off[b]
 *              or NULL.  NULL randomness -> counter-based generator seeded with `seed`.
 *   wav      : float32 out, sequence b at wav_off[b] = spf * mel_off[b], length spf*T[b]
 *              (spf = vt_hift_samples_per_frame: 480 for Chatterbox).
 * Sequences are packed in order: mel_off[b] = sum_{i<b} T[i]. */
int vt_hift_forward(vt_hift* h, const float* mel, const int32_t* T /* HOST */, int B,
                    const float* f0, const float* phase_vec, const float* noise, uint64_t seed,
                    float* wav, void* workspace, int64_t workspace_bytes, void* stream);

/* Debug/inspection taps used by the parity tests: copies an intermediate of the last forward
 * into `out` as float32 [rows, channels] (channel-last); returns rows*channels or <0. */
int64_t vt_hift_read_tap(vt_hift* h, const char* tap, int seq, float* out, int64_t capacity,
                         void* workspace, void* stream);

/* Device-side timing of the last forward (CUDA events recorded on the launch stream).  With
 * profiling enabled every vt_hift_forward records the whole call and each stage's run of
 * ResBlock convolutions (the dominant, tensor-bound kernel class: 72 launches per forward).
 * vt_hift_read_profile waits for the forward to finish and returns the device milliseconds of
 * both, the algorithmic FLOPs (2*C_in*C_out*k per output step) of the ResBlock launches and
 * their count - the numerator/denominator of bench.py's roofline object. */
int vt_hift_set_profiling(vt_hift* h, int enable);
int vt_hift_read_profile(vt_hift* h, double* total_ms, double* resblock_ms, double* resblock_flops,
                         int* resblock_launches);

/* Profiling timeline of the last forward as text, one "section=milliseconds" line per section
 * (zero_gaps, f0_predictor, source_stft, conv_pre, ups<i>, source_down<i>, source_resblock<i>,
 * resblocks<i>, conv_post, istft_head).  `out` is a HOST buffer of `capacity` bytes. */
int vt_hift_read_timeline(vt_hift* h, char* out /* HOST */, int capacity);

/* Number of kernels launched by the last vt_hift_forward / vt_post_process on this thread. */
int vt_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VOCALIE_B200_H */
