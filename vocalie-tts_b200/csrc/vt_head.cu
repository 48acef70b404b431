// Spectral head of the HiFT vocoder: conv_post output -> exp / sin -> inverse STFT (n_fft 16,
// hop 4, periodic Hann, center=True) -> clamp(+-0.99) -> S3Token2Wav trim_fade.
// Upstream hifigan.py HiFTGenerator.decode tail + _istft, s3gen.py S3Token2Wav.inference
// (SURVEY.md Appendix A.5).  Bandwidth-bound: reads 18 floats per frame, writes 4 samples per frame.
//
// torch.istft == per frame irfft (imaginary parts of bins 0 and 8 ignored), times window,
// overlap-add at hop 4, divided by the overlap-added squared window, 8 samples dropped at both ends.
#include "vt_hift.cuh"
#include "vt_tables.cuh"

namespace vt {

constexpr int kHeadSamples = 496;                       // output samples per block (1984 B: whole 32-byte sectors)
constexpr int kHeadFrames = kHeadSamples / kHop + 4;    // frames touching them: 128 = two threads per frame in phase 2

// Three phases per block of 496 output samples:
//   1. (frame, bin) -> re, im   = clip(exp(x_m), 100) * (cos, sin)(sin(x_p))           [shared memory]
//   2. (frame, half) -> 8 of the frame's 16 windowed time samples: real inverse DFT with compile-time twiddles (a
//      data-dependent twiddle index would serialise the constant cache), times hann / 16; all 256 threads work  [shared memory]
//   3. sample -> overlap-add of its <= 4 frames, divide by the overlap-added squared window, clamp, trim_fade
// kFast (the tensor-core operand modes): exp / sin / cos on the SFU.  The arguments are bounded (|sin(x_p)| <= 1; x_m, x_p
// are conv_post outputs, O(10)), so the absolute errors are ~1e-6 of a sample that is clamped to 0.99 - three orders below
// the fp16 operand rounding of the path that feeds it.  The fp32 mode keeps the libm-accurate functions.
// The first version (libm functions everywhere, one thread per frame in phase 2, 64-bit index arithmetic per sample) was
// bound by instruction issue at 2.2 TB/s effective; see DESIGN.md section 4.
template <bool kFast>
__global__ void __launch_bounds__(256)
k_istft_head(const float* __restrict__ post, const int* __restrict__ mel_off, const int* __restrict__ T,
             const long long* __restrict__ off2, const float* __restrict__ trim_fade, int trim_len, int spf,
             float* __restrict__ wav) {
  __shared__ float sh_re[kHeadFrames][10];
  __shared__ float sh_im[kHeadFrames][10];
  __shared__ float sh_y[kHeadFrames][17];
  __shared__ float sh_w2[16];
  const int b = blockIdx.y;
  const long long L = (long long)T[b] * spf;
  const long long frames = L / kHop + 1;
  const long long p0 = (long long)blockIdx.x * kHeadSamples;
  if (p0 >= L) return;
  // first frame overlapping padded position P = p0 + 8:  4f + 15 >= P  ->  f >= (P - 15) / 4 = (p0 - 7) / 4; p0 is a
  // multiple of 16, so the ceiling is p0 / 4 - 1 (0 for the first block)
  const long long f0 = p0 > 0 ? p0 / 4 - 1 : 0;
  const int n_fr = (int)(frames - f0 < kHeadFrames ? frames - f0 : kHeadFrames);   // frames of this block that exist
  const float* src = post + (off2[b] + f0) * kSpecCh;
  if (threadIdx.x < 16) sh_w2[threadIdx.x] = c_hann16[threadIdx.x] * c_hann16[threadIdx.x];
  for (int i = threadIdx.x; i < kHeadFrames * 9; i += blockDim.x) {
    const int fl = i / 9, m = i - fl * 9;
    float re = 0.0f, im = 0.0f;
    if (fl < n_fr) {
      const float xm = src[fl * kSpecCh + m];
      const float xp = src[fl * kSpecCh + 9 + m];
      float mag, sn, cs;
      if constexpr (kFast) {
        mag = fminf(__expf(xm), 100.0f);
        __sincosf(__sinf(xp), &sn, &cs);
      } else {
        mag = fminf(expf(xm), 100.0f);             // torch.clip(mag, max=1e2)
        sincosf(sinf(xp), &sn, &cs);               // phase = sin(x[:, 9:])
      }
      re = mag * cs;
      im = mag * sn;
    }
    sh_re[fl][m] = re;
    sh_im[fl][m] = im;
  }
  __syncthreads();
  {
    const int fl = threadIdx.x >> 1, half = threadIdx.x & 1;
    float re2[8], im2[8];                          // 2 re[m], -2 im[m] for m = 1..7
    const float re0 = sh_re[fl][0], re8 = sh_re[fl][8];
#pragma unroll
    for (int m = 1; m < 8; ++m) { re2[m] = 2.0f * sh_re[fl][m]; im2[m] = -2.0f * sh_im[fl][m]; }
    constexpr float kCos[16] = {1.0f, 0.9238795325112867f, 0.7071067811865476f, 0.3826834323650898f, 0.0f,
                                -0.3826834323650898f, -0.7071067811865476f, -0.9238795325112867f, -1.0f,
                                -0.9238795325112867f, -0.7071067811865476f, -0.3826834323650898f, 0.0f,
                                0.3826834323650898f, 0.7071067811865476f, 0.9238795325112867f};
    constexpr float kHann[16] = {0.0f, 0.03806023374435663f, 0.14644660940672627f, 0.3086582838174551f, 0.5f,
                                 0.6913417161825449f, 0.8535533905932737f, 0.9619397662556434f, 1.0f,
                                 0.9619397662556434f, 0.8535533905932737f, 0.6913417161825449f, 0.5f,
                                 0.3086582838174551f, 0.14644660940672627f, 0.03806023374435663f};
    // both halves run the same unrolled code on their own 8 samples (n = j or 8 + j): the twiddle of (m, 8 + j) is the
    // twiddle of (m, j) times (-1)^m, so the half only flips the sign of the odd bins
    const float sg = half ? -1.0f : 1.0f;
#pragma unroll
    for (int m = 1; m < 8; m += 2) { re2[m] *= sg; im2[m] *= sg; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = re0 + ((j & 1) ? -re8 : re8);    // (-1)^n re8: n and j have the same parity
#pragma unroll
      for (int m = 1; m < 8; ++m) {
        const int ph = (m * j) & 15;
        acc = fmaf(re2[m], kCos[ph], acc);
        acc = fmaf(im2[m], kCos[(ph + 12) & 15], acc);   // sin(t) = cos(t - pi/2)
      }
      const float hn = half ? kHann[8 + j] : kHann[j];
      sh_y[fl][half * 8 + j] = acc * (1.0f / 16.0f) * hn;
    }
  }
  __syncthreads();
  const int n_out = (int)(L - p0 < kHeadSamples ? L - p0 : kHeadSamples);
  const int fdelta = (int)(p0 / 4 - f0);           // local index of frame p0 / 4 (1, or 0 in the first block)
  const int last_fl = n_fr - 1;
  float* out = wav + (long long)mel_off[b] * spf + p0;
  for (int l = threadIdx.x; l < n_out; l += blockDim.x) {
    // padded position P = p0 + l + 8; frames f with 0 <= P - 4 f <= 15, local index f - f0
    const int Pl = l + 8;                          // relative to p0 (a multiple of 4)
    int fb = (Pl >> 2) + fdelta;                   // local index of floor(P / 4)
    int fa = fb - 3;
    int n = (Pl & 3) + 12;                         // P - 4 f at f = fa
    if (fa < 0) { n += 4 * fa; fa = 0; }           // before the sequence's first frame (first block only)
    if (fb > last_fl) fb = last_fl;
    float num = 0.0f, den = 0.0f;
    for (int f = fa; f <= fb; ++f, n -= 4) {
      num += sh_y[f][n];
      den += sh_w2[n];
    }
    float y = num / den;
    y = fminf(fmaxf(y, -0.99f), 0.99f);
    const long long p = p0 + l;
    if (p < trim_len) y *= trim_fade[p];
    out[l] = y;
  }
}

int launch_istft_head(const float* post, const int* mel_off, const int* T, const long long* off2, int B,
                      int T_max, const float* trim_fade, int trim_len, int spf, bool fast, float* wav, cudaStream_t st) {
  if (B == 0 || T_max == 0) return VT_OK;
  const long long Lmax = (long long)T_max * spf;
  dim3 grid((unsigned)((Lmax + kHeadSamples - 1) / kHeadSamples), B);
  if (fast) k_istft_head<true><<<grid, 256, 0, st>>>(post, mel_off, T, off2, trim_fade, trim_len, spf, wav);
  else k_istft_head<false><<<grid, 256, 0, st>>>(post, mel_off, T, off2, trim_fade, trim_len, spf, wav);
  VT_LAUNCHED();
  return VT_OK;
}

}  // namespace vt
