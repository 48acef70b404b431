// Internal structures of the HiFT vocoder path (upstream chatterbox-tts==0.1.6
// chatterbox/models/s3gen/hifigan.py as instantiated by S3Token2Wav).
//
// Data layout in HBM
//   Activations are channel-last: one row per time step, channels contiguous.  The ragged batch
//   is packed along the row axis; at the three generator levels (8T, 40T, 120T+1 rows per
//   sequence) every sequence is preceded and followed by kGap zero rows, so a conv's zero padding
//   is "read the gap" for the TMA-fed tensor-core kernels (the CUDA-core kernels bounds-check
//   explicitly and never rely on it).  Row of step r of sequence b at level l:
//       off_l(b) + r,   off_l(b) = kGap + sum_{b'<b} (len_l(b') + kGap)
//   mel-rate tensors (mel, F0 predictor, conv_pre) are packed without gaps: row mel_off[b] + t.
//   1-D sample-rate signals (source s, wav) are packed without gaps at 480 * mel_off[b].
#pragma once
#include "vt_common.cuh"

#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <string>
#include <vector>

namespace vt {

constexpr int kGap = 32;           // >= largest one-sided conv halo (k=11, dilation 5 -> 25)
constexpr int kMel = 80;
constexpr int kBase = 512;
constexpr int kF0Ch = 512;
constexpr int kHarm = 9;           // nb_harmonics + 1
constexpr int kMaxLevels = 3;      // upsampling stages of the generator (2 or 3)
constexpr int kNfft = 16;
constexpr int kHop = 4;
constexpr int kSpecCh = 32;        // 18 STFT / conv_post channels padded to 32 (zeros)
constexpr int kSpecOp = 24;        // STFT operand rows of the tensor-core source_downs: 18 channels + 6 zeros (48 B)
constexpr int kMelOp = 128;        // mel operand rows: 80 channels + 48 zeros
constexpr int kTileQ = 64;         // output steps per tile of the CUDA-core conv kernel

// Constructor arguments of upstream HiFTGenerator that differ between its users (Chatterbox S3Gen: rates 8/5/3 at 24 kHz,
// tts_backends/chatterbox_impl.py:189; CosyVoice-300M: rates 8/8 at 22.05 kHz, tts_backends/cosyvoice_runner.py:75-131).
struct HiftCfg {
  int n_levels = 3;
  int sr = 24000;
  int up_rate[kMaxLevels] = {8, 5, 3};
  int up_kernel[kMaxLevels] = {16, 11, 7};
  int src_rb_kernel[kMaxLevels] = {7, 7, 11};
  int trim_fade = 1;                  // S3Token2Wav tail (Chatterbox): first sr/50 samples zeroed, next sr/50 faded in
  // derived
  int level_mul[kMaxLevels] = {8, 40, 120};   // steps per mel frame at each level
  int sd_k[kMaxLevels] = {30, 6, 1}, sd_s[kMaxLevels] = {15, 3, 1}, sd_p[kMaxLevels] = {7, 1, 0};   // source_downs
  int spf = 480;                      // samples per mel frame = level_mul[last] * hop
  int trim_n = 480;                   // sr / 50
};

enum ActKind : int { ACT_NONE = 0, ACT_SNAKE = 1, ACT_LRELU = 2, ACT_ELU = 3 };
enum ElemKind : int { ELEM_F32 = 0, ELEM_F16 = 1, ELEM_BF16 = 2 };

// One tile of output steps of one sequence (all conv kernels iterate over a table of these).
struct __align__(16) ConvTile {
  long long in_row0;    // packed row of the sequence's input step 0
  long long out_row0;   // packed row of the sequence's output step 0
  int in_len;           // input steps of the sequence
  int out_len;          // output steps of the sequence ("conv space", before the phase expansion)
  int q0;               // first output step of this tile
  int n;                // output steps in this tile
};

struct ActOut {
  void* dst;            // [rows][C] in the activation element type; nullptr = unused
  const float* alpha;   // per-channel Snake alpha
  int kind;
  float slope;
};

// Arguments of one convolution launch (shared by the CUDA-core and the tensor-core kernels).
struct ConvArgs {
  const float* in;      // fp32 input [rows][in_ld]  (CUDA-core path)
  const void* in_act;   // activation-typed input [rows][in_ld] (resblock convs)
  const void* in_act2;  // second operand source of the K-blocked kernel (low term of a split operand)
  int in_ld;
  const float* w;       // [k][cin][cout] fp32
  const float* bias;    // [cout]
  const float* wscale;  // [cout] tensor-core paths: accumulators hold s_c * (W a) with the exact power of two s_c the host
                        // folded into row c of the packed weights; wscale[c] = 1 / s_c, applied as fma(acc, wscale, bias)
  int cin, cout, k, dil, stride, pad;
  int pro_act;          // prologue activation applied to the input on load
  float pro_slope;
  // epilogue: v = acc + bias (+res1) (+res2); out = (accum ? out : 0) + v*scale; act_i = f_i(v)
  const float* res1;
  const float* res2;
  float* out;
  int out_accum;
  float out_scale;
  ActOut act[3];
  int act_from_out;     // 1: act[0] is computed from the final output value (after scale/accumulate), not from v
  // output row mapping: column c' -> phase r = c'/phase_c, channel c'%phase_c, packed row
  //   out_row0 + q*out_mul + r + out_shift ; dup_row2: value landing on row 2 is also written to row 0
  int out_mul, out_shift, phase_c, dup_row2;
  const ConvTile* tiles;
  int n_tiles;
  unsigned long long tap_skip;   // activation-resident conv: bit (4*column_tile + tap) set = that tap's weights are all zero
  int dbg;              // debug: timing-ablation bits (VT_TC_DBG), 0 normally
  int mc;               // pair kernel: weight ring multicast across a 2-CTA cluster
  long long* trace;     // debug: per-tile role timestamps of CTA 0 (VT_TC_TRACE=<layer>), nullptr normally
};

// One 64-element K block of the K-blocked tensor-core kernel (vt_gemm_tc.cu).
struct __align__(16) KBlock {
  int src;              // operand source buffer: 0 = in_act, 1 = in_act2
  int row_shift;        // input row = in_row0 + q*stride + row_shift
  int ch_off;           // element offset inside the input row (may run past in_ld: im2col of strided convs)
  int reserved;
};

struct ConvLayer {
  std::string name;
  int cin = 0, cout = 0, k = 1, dil = 1, stride = 1, pad = 0;
  int out_mul = 1, phase_c = 0;     // transposed convs run as a k'=3 conv with cout = s*C_out
  float* w = nullptr;               // device [k][cin][cout] fp32
  float* bias = nullptr;            // device [cout]
  float* wscale = nullptr;          // device [cout]: 1 / (power-of-two row scale of the tensor-core weight images); ones if !scaled
  bool scaled = false;              // some row of this layer needed a scale (fp16-subnormal or near-overflow weights)
  void* w_tp = nullptr;             // device, tap-pair image for the C = 64 pair kernel (vt_pair64_tc.cu)
  void* w_tc = nullptr;             // device, tensor-core operand packing (vt_conv_tc.cu)
  unsigned long long tap_skip = 0;  // all-zero (column tile, tap) pairs of a phase-decomposed transposed conv
  void* w_gemm = nullptr;           // device, K-blocked tensor-core packing (vt_gemm_tc.cu)
  void* d_kb = nullptr;             // device KBlock table
  int n_kb = 0, gemm_nt = 0, gemm_elem = 0;
  double flops_per_step = 0;        // algorithmic 2*MAC per output step of the ORIGINAL layer
};

int launch_conv_ref(const ConvArgs& a, int act_elem, cudaStream_t st);
int pack_gemm_tc(ConvLayer& L, const std::vector<KBlock>& kbs, const std::vector<int>& part, const std::vector<float>& wkb,
                 int NT, int elem, std::vector<void*>& allocs);
int launch_gemm_tc(const ConvArgs& a, const ConvLayer& L, int act_elem, cudaStream_t st);

// Source path (vt_source.cu)
int launch_f0_head(const float* h, const float* w, const float* b, float* f0, long long rows, cudaStream_t st);
int launch_sine_source(const float* f0, const int* mel_off, const int* T, int B, long long total_T,
                       const float* phase_vec, const float* noise, unsigned long long seed,
                       const float* lin_w, const float* lin_b, double* phase_base, float* s, int spf, int sr, cudaStream_t st);
int launch_stft(const float* s, const int* mel_off, const int* T, const long long* off2, int B, long long total_T,
                float* spec, void* spec_op, int op_elem, int spf, cudaStream_t st);
// mel fp32 [total_T][80] -> two fp16 terms [gapped rows][128] (hi, lo), channels 80..127 zero
int launch_pack_mel(const float* mel, const int* mel_off, const int* T, const long long* offM, int B, long long total_T,
                    void* mel_hi, void* mel_lo, cudaStream_t st);
// Spectral head (vt_head.cu): conv_post output -> exp/sin -> iSTFT -> clamp -> trim_fade
int launch_istft_head(const float* post, const int* mel_off, const int* T, const long long* off2, int B,
                      int T_max, const float* trim_fade, int trim_len, int spf, bool fast, float* wav, cudaStream_t st);

}  // namespace vt
