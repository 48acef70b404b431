// CUDA-core (fp32 FFMA) convolution over channel-last activations: the exact-arithmetic path of
// the HiFT vocoder.  It runs every layer when the handle is created with VT_OPERAND_FP32, and in
// the tensor-core modes the layers that are not (yet) on tcgen05 (conv_pre, source_downs,
// transposed convs, conv_post, F0 predictor).  Zero padding is an explicit bounds check here.
//
// One launch = one conv layer over a table of (sequence, 64-step) tiles:
//   acc[q][c'] = sum_j sum_ci f(in[q*stride + j*dil - pad][ci]) * w[j][ci][c']
// followed by the shared epilogue (bias, up to two residual streams, scaled/accumulated fp32
// output, up to three activated copies for the next layers).  Transposed convs arrive here as a
// 3-tap conv with cout = stride*C_out ("phase-decomposed"), the epilogue scatters phase r of step
// q to output row q*stride + r.
#include "vt_hift.cuh"

namespace vt {

namespace {

template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void ld4(const float* p, float v[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void st4(float* p, const float v[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Elem<__half> {
  static __device__ __forceinline__ void ld4(const __half* p, float v[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void st4(__half* p, const float v[4]) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<unsigned*>(&a);
    t.y = *reinterpret_cast<unsigned*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ void ld4(const __nv_bfloat16* p, float v[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void st4(__nv_bfloat16* p, const float v[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<unsigned*>(&a);
    t.y = *reinterpret_cast<unsigned*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

}  // namespace

// Activation functions of the path (fp32).  Snake: upstream hifigan.py Snake.forward,
// x + 1/(alpha + 1e-9) * sin(alpha*x)^2.
__device__ __forceinline__ float apply_act(float v, int kind, float alpha, float slope) {
  switch (kind) {
    case ACT_SNAKE: {
      const float s = sinf(v * alpha);
      return v + (1.0f / (alpha + 1e-9f)) * (s * s);
    }
    case ACT_LRELU: return v > 0.0f ? v : v * slope;
    case ACT_ELU: return v > 0.0f ? v : expm1f(v);
    default: return v;
  }
}

template <typename ActT>
__global__ void __launch_bounds__(256)
k_conv_ref(const ConvArgs a) {
  __shared__ __align__(16) float As[16][kTileQ + 4];
  __shared__ __align__(16) float Ws[16][64];
  const ConvTile tile = a.tiles[blockIdx.x];
  const int c0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.0f;

  const int a_row = tid >> 2, a_cg = (tid & 3) * 4;   // A loader: 64 rows x 16 channels
  const int w_ci = tid >> 4, w_cg = (tid & 15) * 4;   // W loader: 16 channels x 64 columns
  const int q_ld = tile.q0 + a_row;
  const ActT* in_act = reinterpret_cast<const ActT*>(a.in_act);

  for (int j = 0; j < a.k; ++j) {
    const int r = q_ld * a.stride + j * a.dil - a.pad;
    const bool row_ok = (a_row < tile.n) && (r >= 0) && (r < tile.in_len);
    const long long in_base = (tile.in_row0 + r) * (long long)a.in_ld;
    for (int ci0 = 0; ci0 < a.cin; ci0 += 16) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (row_ok) {
        if (in_act) Elem<ActT>::ld4(in_act + in_base + ci0 + a_cg, v);
        else Elem<float>::ld4(a.in + in_base + ci0 + a_cg, v);
        if (a.pro_act != ACT_NONE) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = apply_act(v[e], a.pro_act, 1.0f, a.pro_slope);
        }
      }
      float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + w_cg < a.cout)
        wv = *reinterpret_cast<const float4*>(a.w + ((long long)(j * a.cin + ci0 + w_ci)) * a.cout + c0 + w_cg);
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) As[a_cg + e][a_row] = v[e];
      *reinterpret_cast<float4*>(&Ws[w_ci][w_cg]) = wv;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
        const float ar[4] = {av.x, av.y, av.z, av.w};
        const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][e] = fmaf(ar[i], br[e], acc[i][e]);
      }
    }
  }

  // ---- epilogue
  const int cc = c0 + tx * 4;
  if (cc >= a.cout) return;
  const int phase = cc / a.phase_c;
  const int co = cc - phase * a.phase_c;
  float bias[4];
  Elem<float>::ld4(a.bias + cc, bias);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ql = ty * 4 + i;
    if (ql >= tile.n) continue;
    const long long step = (long long)(tile.q0 + ql) * a.out_mul + phase + a.out_shift;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = acc[i][e] + bias[e];
    for (int rep = 0; rep < 2; ++rep) {
      long long orow = tile.out_row0 + step;
      if (rep == 1) {
        if (!(a.dup_row2 && step == 2)) break;
        orow = tile.out_row0;  // reflection pad: padded[0] = unpadded[1]
      }
      const long long idx = orow * a.phase_c + co;
      float t[4] = {v[0], v[1], v[2], v[3]};
      if (a.res1) {
        float r1[4];
        Elem<float>::ld4(a.res1 + idx, r1);
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] += r1[e];
      }
      if (a.res2) {
        float r2[4];
        Elem<float>::ld4(a.res2 + idx, r2);
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] += r2[e];
      }
      float o[4] = {t[0], t[1], t[2], t[3]};
      if (a.out) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = t[e] * a.out_scale;
        if (a.out_accum) {
          float p[4];
          Elem<float>::ld4(a.out + idx, p);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] += p[e];
        }
        Elem<float>::st4(a.out + idx, o);
      }
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (!a.act[s].dst) continue;
        float al[4] = {1.f, 1.f, 1.f, 1.f};
        if (a.act[s].alpha) Elem<float>::ld4(a.act[s].alpha + co, al);
        const float* srcv = (s == 0 && a.act_from_out) ? o : t;
        float y[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) y[e] = apply_act(srcv[e], a.act[s].kind, al[e], a.act[s].slope);
        Elem<ActT>::st4(reinterpret_cast<ActT*>(a.act[s].dst) + idx, y);
      }
    }
  }
}

int launch_conv_ref(const ConvArgs& a, int act_elem, cudaStream_t st) {
  VT_REQUIRE(a.cin % 16 == 0 && a.cout % 4 == 0 && a.phase_c % 4 == 0 && a.in_ld % 4 == 0,
             "conv_ref: cin must be a multiple of 16, cout/phase_c/in_ld multiples of 4 (cin=%d cout=%d)", a.cin, a.cout);
  if (a.n_tiles == 0) return VT_OK;
  dim3 grid(a.n_tiles, (a.cout + 63) / 64);
  switch (act_elem) {
    case ELEM_F32: k_conv_ref<float><<<grid, 256, 0, st>>>(a); break;
    case ELEM_F16: k_conv_ref<__half><<<grid, 256, 0, st>>>(a); break;
    case ELEM_BF16: k_conv_ref<__nv_bfloat16><<<grid, 256, 0, st>>>(a); break;
    default: VT_REQUIRE(false, "conv_ref: bad activation element kind %d", act_elem);
  }
  VT_LAUNCHED();
  return VT_OK;
}

}  // namespace vt
