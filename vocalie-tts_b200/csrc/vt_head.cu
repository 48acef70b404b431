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

constexpr int kHeadSamples = 512;                       // output samples per block
constexpr int kHeadFrames = kHeadSamples / kHop + 4;    // frames touching them (132)

// Three phases per block of 512 output samples:
//   1. (frame, bin) -> re, im   = clip(exp(x_m), 100) * (cos, sin)(sin(x_p))           [shared memory]
//   2. frame -> 16 windowed time samples: real inverse DFT with compile-time twiddles (one thread per frame;
//      a data-dependent twiddle index would serialise the constant cache), times hann / 16      [shared memory]
//   3. sample -> overlap-add of its <= 4 frames, divide by the overlap-added squared window, clamp, trim_fade
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
  // first frame overlapping padded position P = p0 + 8:  4f + 15 >= P  ->  f >= (P - 15) / 4
  long long f0 = (p0 + 8 - 15 + 3) / 4;   // ceil((p0 - 7) / 4) for p0 >= 0 (p0 multiple of 512)
  if (f0 < 0) f0 = 0;
  const float* src = post + (off2[b] + f0) * kSpecCh;
  if (threadIdx.x < 16) sh_w2[threadIdx.x] = c_hann16[threadIdx.x] * c_hann16[threadIdx.x];
  for (int i = threadIdx.x; i < kHeadFrames * 9; i += blockDim.x) {
    const int fl = i / 9, m = i - fl * 9;
    float re = 0.0f, im = 0.0f;
    if (f0 + fl < frames) {
      const float xm = src[(long long)fl * kSpecCh + m];
      const float xp = src[(long long)fl * kSpecCh + 9 + m];
      const float mag = fminf(expf(xm), 100.0f);   // torch.clip(mag, max=1e2)
      const float ph = sinf(xp);                   // phase = sin(x[:, 9:])
      float sn, cs;
      sincosf(ph, &sn, &cs);
      re = mag * cs;
      im = mag * sn;
    }
    sh_re[fl][m] = re;
    sh_im[fl][m] = im;
  }
  __syncthreads();
  if (threadIdx.x < kHeadFrames) {
    const int fl = threadIdx.x;
    float re[9], im[9];
#pragma unroll
    for (int m = 0; m < 9; ++m) { re[m] = sh_re[fl][m]; im[m] = sh_im[fl][m]; }
    constexpr float kCos[16] = {1.0f, 0.9238795325112867f, 0.7071067811865476f, 0.3826834323650898f, 0.0f,
                                -0.3826834323650898f, -0.7071067811865476f, -0.9238795325112867f, -1.0f,
                                -0.9238795325112867f, -0.7071067811865476f, -0.3826834323650898f, 0.0f,
                                0.3826834323650898f, 0.7071067811865476f, 0.9238795325112867f};
    constexpr float kHann[16] = {0.0f, 0.03806023374435663f, 0.14644660940672627f, 0.3086582838174551f, 0.5f,
                                 0.6913417161825449f, 0.8535533905932737f, 0.9619397662556434f, 1.0f,
                                 0.9619397662556434f, 0.8535533905932737f, 0.6913417161825449f, 0.5f,
                                 0.3086582838174551f, 0.14644660940672627f, 0.03806023374435663f};
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      float acc = re[0] + ((n & 1) ? -re[8] : re[8]);
#pragma unroll
      for (int m = 1; m < 8; ++m) {
        const int ph = (m * n) & 15;
        acc = fmaf(2.0f * re[m], kCos[ph], acc);
        acc = fmaf(-2.0f * im[m], kCos[(ph + 12) & 15], acc);   // sin(t) = cos(t - pi/2)
      }
      sh_y[fl][n] = acc * (1.0f / 16.0f) * kHann[n];
    }
  }
  __syncthreads();
  for (int l = threadIdx.x; l < kHeadSamples; l += blockDim.x) {
    const long long p = p0 + l;
    if (p >= L) break;
    const long long P = p + 8;
    long long fa = (P - 15 + 3) >> 2;   // ceil((P-15)/4); P-15+3 >= -4 -> arithmetic shift is a floor
    if (fa < 0) fa = 0;
    long long fb = P >> 2;
    if (fb > frames - 1) fb = frames - 1;
    float num = 0.0f, den = 0.0f;
    for (long long f = fa; f <= fb; ++f) {
      const int n = (int)(P - 4 * f);          // 0..15
      num += sh_y[(int)(f - f0)][n];
      den += sh_w2[n];
    }
    float y = num / den;
    y = fminf(fmaxf(y, -0.99f), 0.99f);
    if (p < trim_len) y *= trim_fade[p];
    wav[(long long)mel_off[b] * spf + p] = y;
  }
}

int launch_istft_head(const float* post, const int* mel_off, const int* T, const long long* off2, int B,
                      int T_max, const float* trim_fade, int trim_len, int spf, float* wav, cudaStream_t st) {
  if (B == 0 || T_max == 0) return VT_OK;
  const long long Lmax = (long long)T_max * spf;
  dim3 grid((unsigned)((Lmax + kHeadSamples - 1) / kHeadSamples), B);
  k_istft_head<<<grid, 256, 0, st>>>(post, mel_off, T, off2, trim_fade, trim_len, spf, wav);
  VT_LAUNCHED();
  return VT_OK;
}

}  // namespace vt
