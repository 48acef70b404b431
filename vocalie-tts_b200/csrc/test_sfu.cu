// Microbenchmark (GPU box, debugging sessions): throughput of the Snake activation per SM as a function of the
// number of resident warps, for sin.approx on the SFU (snake_f) and the FMA-pipe polynomial (snake_poly_f).
// Prints elements per clock per SM.  Usage: build/test_sfu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float snake_mufu(float v, float alpha, float inv_alpha) {
  const float s = __sinf(v * alpha);
  return fmaf(inv_alpha, s * s, v);
}
// u = v*alpha/pi, r = u - rint(u) in [-0.5, 0.5]; sin^2(pi r) = t*P(t), t = r^2
__device__ __forceinline__ float snake_poly(float v, float a_pi, float inv_alpha) {
  const float m = fmaf(v, a_pi, 12582912.0f);
  const float rn = m - 12582912.0f;
  const float r = fmaf(v, a_pi, -rn);
  const float t = r * r;
  float p = fmaf(t, 10.603050f, -29.434399f);
  p = fmaf(t, p, 42.643887f);
  p = fmaf(t, p, -32.465050f);
  p = fmaf(t, p, 9.8695183f);
  return fmaf(inv_alpha * t, p, v);
}

template <int MODE>   // 0: all MUFU, 1: all poly, 2: 4 MUFU + 4 poly per group of 8, 3: 2 MUFU + 6 poly
__global__ void k_bench(float* out, int iters, long long* cycles) {
  float x[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] = 0.001f * (threadIdx.x + e * 37);
  const float alpha = 1.0f + 0.01f * (threadIdx.x & 7), ia = 1.0f / alpha, a_pi = alpha * 0.318309886f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const bool poly = MODE == 1 || (MODE == 2 && (e & 1)) || (MODE == 3 && (e & 3));
      x[e] = poly ? snake_poly(x[e], a_pi, ia) : snake_mufu(x[e], alpha, ia);
      x[e] = x[e] * 0.5f;   // keep the values bounded
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) s += x[e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const char* names[4] = {"mufu", "poly", "4+4", "2+6"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps = 1; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k_bench<0><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 1) k_bench<1><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 2) k_bench<2><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 3) k_bench<3><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      long long h = 0;
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%s warps=%2d: %.2f snakes/clk/SM (%.1f cycles per warp-snake)\n", names[mode], warps,
             (double)iters * 8 * warps * 32 / (double)h, (double)h / (iters * 8));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
