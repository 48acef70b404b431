// Source path of the HiFT vocoder: F0 head, SineGen + SourceModuleHnNSF, STFT of the source.
// Bandwidth-bound elementwise work over sample-rate signals (upstream hifigan.py SineGen.forward,
// SourceModuleHnNSF.forward, HiFTGenerator._stft).
#include "vt_hift.cuh"
#include "vt_tables.cuh"

namespace vt {

// ---- F0 head: classifier Linear(512 -> 1) + abs (upstream f0_predictor.py) -----------------------
__global__ void __launch_bounds__(256)
k_f0_head(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ b,
          float* __restrict__ f0, long long rows) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* hp = reinterpret_cast<const float4*>(h + row * kF0Ch);
  const float4* wp = reinterpret_cast<const float4*>(w);
  float acc = 0.0f;
#pragma unroll
  for (int i = 0; i < kF0Ch / 128; ++i) {
    const float4 x = hp[i * 32 + lane], y = wp[i * 32 + lane];
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
    acc = fmaf(x.z, y.z, acc);
    acc = fmaf(x.w, y.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) f0[row] = fabsf(acc + b[0]);
}

int launch_f0_head(const float* h, const float* w, const float* b, float* f0, long long rows, cudaStream_t st) {
  if (rows == 0) return VT_OK;
  k_f0_head<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(h, w, b, f0, rows);
  VT_LAUNCHED();
  return VT_OK;
}

// ---- SineGen -------------------------------------------------------------------------------
// Phase: torch CPU cumsum over fp32 accumulates in fp64 and rounds every prefix to fp32
// (SURVEY A.4 [probe]).  F is constant inside a mel frame, so the fp64 prefix at sample j of
// frame t is base[t] + (j+1)*F with base[t] = sum_{t'<t} 480*F(t') (exact products in fp64);
// it is rounded to fp32 *before* the mod-1 exactly like the reference.
__device__ __forceinline__ float harmonic_inc(float f0, int h, float sr) {
  // upstream: F_mat = f0 * (i + 1) / sampling_rate   (fp32 mul, then fp32 div)
  return __fdiv_rn(__fmul_rn(f0, (float)(h + 1)), sr);
}

// One WARP per (sequence, harmonic): every lane sums a contiguous run of frames, the lane totals are scanned with
// shuffles, then the lane writes the prefixes of its run.  (One THREAD per (sequence, harmonic) walked the 500 frames of
// a 10 s chunk serially: 78 us of latency for 576 threads' worth of work.)  The terms spf * F are exact in fp64 (9 + 24
// significant bits) and so are their partial sums as long as the terms' exponents lie within ~2^10 of each other, so the
// order of the additions does not change the prefixes.
__global__ void __launch_bounds__(128)
k_phase_base(const float* __restrict__ f0, const int* __restrict__ mel_off,
             const int* __restrict__ T, int B, long long total_T, double* __restrict__ base, int spf, float sr) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * kHarm) return;
  const int b = w / kHarm, h = w % kHarm;
  const long long o = mel_off[b];
  const int Tb = T[b];
  const int run = (Tb + 31) / 32;
  const int t0 = lane * run < Tb ? lane * run : Tb, t1 = t0 + run < Tb ? t0 + run : Tb;
  double local = 0.0;
  for (int t = t0; t < t1; ++t) local += (double)spf * (double)harmonic_inc(f0[o + t], h, sr);
  double incl = local;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  double acc = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) acc = 0.0;
  double* dst = base + (long long)h * total_T + o;
  double* inc = dst + (long long)kHarm * total_T;      // second half of the buffer: the per-frame increment itself
  for (int t = t0; t < t1; ++t) {
    const double d = (double)harmonic_inc(f0[o + t], h, sr);
    dst[t] = acc;
    inc[t] = d;                                          // (k_sine_source: no IEEE division per sample and harmonic)
    acc += (double)spf * d;
  }
}

// Philox4x32-7 counter-based generator for the performance mode (no explicit noise given; 7 rounds is the
// fewest that passes BigCrush in Salmon et al., SC'11 - the generator is a third of this kernel's instructions).
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b) {
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
  const float u2 = (float)b * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// sin(t) for t in [-pi, 3pi]: one explicit reduction to [-pi, pi], then the SFU sine (absolute error
// ~5e-7 there: three orders below the fp16 operand rounding of everything the source feeds).
__device__ __forceinline__ float sin_2pi_range(float t) {
  t = fmaf(-6.283185307179586f, rintf(t * 0.15915494309189535f), t);
  return __sinf(t);
}

// One thread per FOUR consecutive output samples (480 samples per frame: the four share the frame, hence f0, the
// voicing decision and the per-harmonic phase prefix / increment): 9 harmonics, noise mix, Linear(9->1), tanh.
// Performance-mode noise: one Philox call per (sample quad, harmonic) yields the quad's four normal deviates.
__global__ void __launch_bounds__(256)
k_sine_source(const float* __restrict__ f0, const int* __restrict__ mel_off, const int* __restrict__ T, int B,
              long long total_T, const float* __restrict__ phase_vec, const float* __restrict__ noise,
              unsigned long long seed, const float* __restrict__ lin_w, const float* __restrict__ lin_b,
              const double* __restrict__ base, float* __restrict__ s, int spf) {
  const int b = blockIdx.y;
  const long long o = mel_off[b];
  const long long L = (long long)T[b] * spf;
  const float lb = lin_b[0];
  float lw[kHarm], pv[kHarm];
#pragma unroll
  for (int h = 0; h < kHarm; ++h) {
    lw[h] = lin_w[h];
    pv[h] = phase_vec ? phase_vec[b * kHarm + h] : 0.0f;
  }
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  if (!phase_vec) {
    // U(-pi, pi) per (sequence, harmonic), harmonic 0 -> 0 (upstream SineGen)
    const uint4 r0 = philox4x32(make_uint4((unsigned)b, 0u, 0u, 0x51u), key);
    const uint4 r1 = philox4x32(make_uint4((unsigned)b, 1u, 0u, 0x51u), key);
    const unsigned rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int h = 1; h < kHarm; ++h) pv[h] = ((float)rr[h - 1] * 2.3283064365386963e-10f * 2.0f - 1.0f) * 3.14159265358979f;
  }
  for (long long n = 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); n < L; n += 4 * (long long)gridDim.x * blockDim.x) {
    const int t = (int)(n / spf), j = (int)(n - (long long)t * spf);
    const float f = f0[o + t];
    const bool voiced = f > 10.0f;
    const float uv = voiced ? 1.0f : 0.0f;
    // upstream: noise_amp = uv * noise_std + (1 - uv) * sine_amp / 3
    // (uv in {0, 1}: both values are exact compile-time constants of the same fp32 expression)
    const float namp = voiced ? 0.003f : 0.1f / 3.0f;
    const unsigned long long gi = (unsigned long long)(o * spf + n) >> 2;      // global index of the quad
    float acc[4] = {lb, lb, lb, lb};
#pragma unroll
    for (int h = 0; h < kHarm; ++h) {
      float z[4];
      if (noise) {
        const float4 zz = *reinterpret_cast<const float4*>(noise + (o * spf) * kHarm + (long long)h * L + n);
        z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
      } else {
        const uint4 r = philox4x32(make_uint4((unsigned)gi, (unsigned)(gi >> 32), (unsigned)h, 0xA5u), key);
        const float2 n0 = box_muller(r.x, r.y), n1 = box_muller(r.z, r.w);
        z[0] = n0.x; z[1] = n0.y; z[2] = n1.x; z[3] = n1.y;
      }
      const long long bi = (long long)h * total_T + o + t;
      const double b64 = base[bi], i64 = base[bi + (long long)kHarm * total_T];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double c64 = b64 + (double)(j + q + 1) * i64;
        const float c = __double2float_rn(c64);
        const float frac = c - truncf(c);                       // torch `% 1` on a non-negative value
        const float theta = __fmul_rn(frac, 6.283185307179586f);  // 2*pi as an fp32 scalar
        const float sine = __fmul_rn(0.1f, sin_2pi_range(theta + pv[h]));
        const float v = __fadd_rn(__fmul_rn(sine, uv), __fmul_rn(namp, z[q]));
        acc[q] = fmaf(v, lw[h], acc[q]);
      }
    }
    *reinterpret_cast<float4*>(s + o * spf + n) = make_float4(tanhf(acc[0]), tanhf(acc[1]), tanhf(acc[2]), tanhf(acc[3]));
  }
}

int launch_sine_source(const float* f0, const int* mel_off, const int* T, int B, long long total_T,
                       const float* phase_vec, const float* noise, unsigned long long seed,
                       const float* lin_w, const float* lin_b, double* phase_base, float* s, int spf, int sr, cudaStream_t st) {
  if (B == 0 || total_T == 0) return VT_OK;
  VT_REQUIRE(spf % 4 == 0, "samples per frame must be a multiple of 4 (a sample quad must not straddle a frame)");
  k_phase_base<<<(B * kHarm + 3) / 4, 128, 0, st>>>(f0, mel_off, T, B, total_T, phase_base, spf, (float)sr);
  VT_LAUNCHED();
  // grid.x sized for the longest sequence (grid-stride inside, four samples per thread): aim at ~148*8 blocks in total
  int gx = (148 * 8 + B - 1) / B;
  if (gx < 1) gx = 1;
  dim3 grid(gx, B);
  k_sine_source<<<grid, 256, 0, st>>>(f0, mel_off, T, B, total_T, phase_vec, noise, seed, lin_w, lin_b,
                                      phase_base, s, spf);
  VT_LAUNCHED();
  return VT_OK;
}

// ---- STFT of the source: n_fft 16, hop 4, periodic Hann, center=True (reflect) ---------------
// One thread per frame: 16 windowed samples -> 9 real + 9 imaginary bins, written channel-last
// into the level-2 packed layout (row off2[b] + frame, 32 channels, 18..31 zero).

template <typename OpT>
__device__ __forceinline__ unsigned pack2_op(float a, float b);
template <> __device__ __forceinline__ unsigned pack2_op<__half>(float a, float b) {
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const unsigned*>(&v);
}
template <> __device__ __forceinline__ unsigned pack2_op<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const unsigned*>(&v);
}

// `spec_op` (tensor-core modes): the same 18 channels as operand-typed rows of kSpecOp = 24 elements.
template <typename OpT>
__global__ void __launch_bounds__(256)
k_stft(const float* __restrict__ s, const int* __restrict__ mel_off, const int* __restrict__ T,
       const long long* __restrict__ off2, int B, float* __restrict__ spec, OpT* __restrict__ spec_op, int spf) {
  const int b = blockIdx.y;
  const long long L = (long long)T[b] * spf;
  const long long frames = L / kHop + 1;
  const float* sb = s + (long long)mel_off[b] * spf;
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < frames; f += (long long)gridDim.x * blockDim.x) {
    float x[kNfft];
#pragma unroll
    for (int n = 0; n < kNfft; ++n) {
      long long i = f * kHop + n - kNfft / 2;
      if (i < 0) i = -i;
      if (i >= L) i = 2 * (L - 1) - i;
      x[n] = sb[i] * c_hann16[n];
    }
    // real DFT through the even / odd parts of the frame: cos(2 pi m (16 - n) / 16) = cos(2 pi m n / 16), sin changes
    // sign, so with e[n] = x[n] + x[16 - n], o[n] = x[n] - x[16 - n] (n = 1..7) the 9 bins take 112 instead of 288 FMAs
    float ev[8], od[8];
#pragma unroll
    for (int n = 1; n < 8; ++n) { ev[n] = x[n] + x[kNfft - n]; od[n] = x[n] - x[kNfft - n]; }
    float outv[kSpecCh];
#pragma unroll
    for (int m = 0; m <= kNfft / 2; ++m) {
      float re = x[0] + ((m & 1) ? -x[8] : x[8]), im = 0.0f;
#pragma unroll
      for (int n = 1; n < 8; ++n) {
        const int ph = (m * n) & 15;
        re = fmaf(ev[n], c_cos16[ph], re);
        if (m > 0 && m < kNfft / 2) im = fmaf(od[n], -c_sin16[ph], im);
      }
      outv[m] = re;
      outv[kNfft / 2 + 1 + m] = im;
    }
#pragma unroll
    for (int c = kNfft + 2; c < kSpecCh; ++c) outv[c] = 0.0f;
    if (spec) {
      float* dst = spec + (off2[b] + f) * kSpecCh;
#pragma unroll
      for (int c = 0; c < kSpecCh; c += 4)
        *reinterpret_cast<float4*>(dst + c) = make_float4(outv[c], outv[c + 1], outv[c + 2], outv[c + 3]);
    }
    if (spec_op) {
      uint4* od = reinterpret_cast<uint4*>(spec_op + (off2[b] + f) * kSpecOp);
#pragma unroll
      for (int c = 0; c < kSpecOp; c += 8)
        od[c / 8] = make_uint4(pack2_op<OpT>(outv[c], outv[c + 1]), pack2_op<OpT>(outv[c + 2], outv[c + 3]),
                               pack2_op<OpT>(outv[c + 4], outv[c + 5]), pack2_op<OpT>(outv[c + 6], outv[c + 7]));
    }
  }
}

int launch_stft(const float* s, const int* mel_off, const int* T, const long long* off2, int B, long long total_T,
                float* spec, void* spec_op, int op_elem, int spf, cudaStream_t st) {
  if (B == 0 || total_T == 0) return VT_OK;
  int gx = (148 * 8 + B - 1) / B;
  dim3 grid(gx < 1 ? 1 : gx, B);
  if (spec_op && op_elem == ELEM_BF16)
    k_stft<__nv_bfloat16><<<grid, 256, 0, st>>>(s, mel_off, T, off2, B, spec, reinterpret_cast<__nv_bfloat16*>(spec_op), spf);
  else
    k_stft<__half><<<grid, 256, 0, st>>>(s, mel_off, T, off2, B, spec, reinterpret_cast<__half*>(spec_op), spf);
  VT_LAUNCHED();
  return VT_OK;
}

// ---- mel operand pack: fp32 [total_T][80] -> fp16 hi / lo terms in the gapped mel-rate layout ----------
__global__ void __launch_bounds__(256)
k_pack_mel(const float* __restrict__ mel, const int* __restrict__ mel_off, const int* __restrict__ T,
           const long long* __restrict__ offM, int B, __half* __restrict__ hi, __half* __restrict__ lo) {
  const int b = blockIdx.y;
  const long long n = (long long)T[b] * (kMelOp / 4);
  const float* src = mel + (long long)mel_off[b] * kMel;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / (kMelOp / 4);
    const int c = (int)(i - t * (kMelOp / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < kMel) v = *reinterpret_cast<const float4*>(src + t * kMel + c);
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    const long long o = (offM[b] + t) * kMelOp + c;
    *reinterpret_cast<uint2*>(hi + o) = make_uint2(*reinterpret_cast<const unsigned*>(&h0), *reinterpret_cast<const unsigned*>(&h1));
    *reinterpret_cast<uint2*>(lo + o) = make_uint2(*reinterpret_cast<const unsigned*>(&l0), *reinterpret_cast<const unsigned*>(&l1));
  }
}

int launch_pack_mel(const float* mel, const int* mel_off, const int* T, const long long* offM, int B, long long total_T,
                    void* mel_hi, void* mel_lo, cudaStream_t st) {
  if (B == 0 || total_T == 0) return VT_OK;
  int gx = (148 * 4 + B - 1) / B;
  dim3 grid(gx < 1 ? 1 : gx, B);
  k_pack_mel<<<grid, 256, 0, st>>>(mel, mel_off, T, offM, B, reinterpret_cast<__half*>(mel_hi), reinterpret_cast<__half*>(mel_lo));
  VT_LAUNCHED();
  return VT_OK;
}

}  // namespace vt
