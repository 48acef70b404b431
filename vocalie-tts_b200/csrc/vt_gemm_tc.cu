// K-blocked tcgen05 GEMM-convolution for the layers around the ResBlocks (F0 predictor trunk,
// conv_pre, source_downs): 4.5 % of the FLOPs, but 30 % of the forward while they ran on CUDA cores.
//
//   out[q][n] = sum_kb  A_kb[q][0..63] . W_kb[0..63][n]
//   A_kb[q][e] = src[kb.src][ (in_row0 + q*stride + kb.row_shift) * in_ld + kb.ch_off + e ]
//
// A "K block" is 64 contiguous operand elements (128 bytes) of one input row; the list of K blocks of a
// layer is a small table built on the host, so one kernel covers
//   * ordinary convolutions      : one K block per (tap, 64-channel block), row_shift = tap*dil - pad;
//   * strided source_downs       : the STFT operand rows are 24 halves (48 B) wide, so the k*24 values a
//                                  strided output step reads are CONTIGUOUS in memory: the layer is a plain
//                                  GEMM whose A rows overlap (row stride 15*24 or 3*24 elements), K blocks
//                                  walk along that run (im2col without materialising anything);
//   * split-precision operands   : x = hi + lo (two fp16 terms, ~22 mantissa bits).  x.w ~ hi.w_hi + hi.w_lo
//                                  + lo.w_hi is three K-block groups reading two source buffers; the F0
//                                  trunk needs it because F0 feeds a 10-second phase integral.
// Pipeline: ST stages, each one K block of A (MB*128 rows x 128 B, SWIZZLE_128B, cp.async by 4 producer
// warps with lagged completion) and of W (NT x 128 B, one bulk-TMA copy); MMA warp; NEPI epilogue warps
// on double-buffered TMEM accumulators (shared epilogue of vt_tc.cuh).
#include "vt_tc.cuh"

#include <cstring>

namespace vt {

namespace tc {

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT, int MB, int ST, int NEPI, int EM, typename ActT>
__global__ void __launch_bounds__((NEPI + 2 + kProdWarps) * 32, 1)
k_gemm_tc(const ConvArgs a, const uint8_t* __restrict__ wtc, const KBlock* __restrict__ kbs, const int n_kb,
          const uint32_t idesc) {
  static_assert(kSwz, "the K-blocked kernel assumes the SWIZZLE_128B operand layout");
  constexpr int W_MMA = NEPI, W_WP = NEPI + 1, W_AP = NEPI + 2;
  constexpr int A_BYTES = MB * 128 * 128, W_BYTES = NT * 128, STAGE = A_BYTES + W_BYTES;
  constexpr int ACC_COLS = MB * NT, TMEM_COLS = pow2_cols(2 * ACC_COLS);
  constexpr int LAG = ST - 2;                      // cp.async groups a producer thread keeps in flight
  static_assert(2 * ACC_COLS <= 512 && ST >= 3 && NT % 32 == 0, "bad tile configuration");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * STAGE);
  uint64_t* full = bars;
  uint64_t* empty = full + ST;
  uint64_t* acc_full = empty + ST;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* stage_all = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tmem_slot) + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) { mbar_init(&full[i], kProdWarps + 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NEPI); }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();                             // dependents may be scheduled (they take an SM when its CTA of this grid exits)
  if (warp != W_WP) pdl_wait();                // everything but the (static) weight stream waits for the previous kernel

  const int n_nt = a.cout / NT;
  const int n_tiles = a.n_tiles * n_nt;            // (row tile, column tile), column tile fastest

  if (warp >= W_AP) {
    // ---------------- A producers: 8 consecutive threads copy one 128-byte row segment
    const int pt = threadIdx.x - W_AP * 32;
    const int c = pt & 7, rbase = pt >> 3;
    const uint32_t swz = (uint32_t)((c ^ (rbase & 7)) << 4);     // (16*i + rbase) & 7 == rbase & 7
    const uint8_t* srcs[2] = {reinterpret_cast<const uint8_t*>(a.in_act), reinterpret_cast<const uint8_t*>(a.in_act2)};
    uint32_t issued = 0, arrived = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const ConvTile tile = a.tiles[t / n_nt];
      for (int kb = 0; kb < n_kb; ++kb) {
        const uint32_t s = issued % ST, ph = (issued / ST) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        const KBlock k = kbs[kb];
        const uint8_t* base = srcs[k.src];
        const uint32_t dst = smem_u32(smem + s * STAGE) + swz;
#pragma unroll
        for (int i = 0; i < 8 * MB; ++i) {
          const int r = i * 16 + rbase;
          const int rr = r < tile.n ? r : tile.n - 1;              // rows past the tile repeat its last row (never stored)
          const long long row = tile.in_row0 + (long long)(tile.q0 + rr) * a.stride + k.row_shift;
          cp_async16(dst + (uint32_t)r * 128u, base + ((row * a.in_ld + k.ch_off + c * 8) << 1));
        }
        cp_async_commit();
        ++issued;
        if (issued - arrived > (uint32_t)LAG) {
          cp_async_wait_group<LAG>();
          fence_proxy_async();
          mbar_arrive_warp(&full[arrived % ST]);
          ++arrived;
        }
      }
    }
    cp_async_wait_all();
    fence_proxy_async();
    for (; arrived < issued; ++arrived) mbar_arrive_warp(&full[arrived % ST]);
  } else if (warp == W_WP) {
    // ---------------- weight producer: one bulk copy per K block
    if (lane == 0) {
      uint32_t issued = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint8_t* wsrc = wtc + (size_t)(t % n_nt) * n_kb * W_BYTES;
        for (int kb = 0; kb < n_kb; ++kb, ++issued) {
          const uint32_t s = issued % ST, ph = (issued / ST) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], W_BYTES);
          bulk_g2s(smem + s * STAGE + A_BYTES, wsrc + (size_t)kb * W_BYTES, W_BYTES, &full[s]);
        }
      }
    }
  } else if (warp == W_MMA) {
    // ---------------- MMA issuer
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lo0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
    uint32_t done = 0;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t asph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&acc_empty[as], asph ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(as * ACC_COLS);
      for (int kb = 0; kb < n_kb; ++kb, ++done) {
        const uint32_t s = done % ST, ph = (done / ST) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = lo0 + s * (uint32_t)(STAGE >> 4);
          const uint32_t b_lo = a_lo + (uint32_t)(A_BYTES >> 4);
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_lh(d0 + (uint32_t)(mb * NT), a_lo + (uint32_t)(mb * 1024 + ks * 2), b_lo + (uint32_t)(ks * 2), kDescHi,
                          idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_full[as]);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue warps
    float* stage = stage_all + warp * (32 * kStageLd);
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t asph = (uint32_t)(it >> 1) & 1u;
      const int rt = t / n_nt, nt = t - rt * n_nt;
      const ConvTile tile = a.tiles[rt];
      mbar_wait(&acc_full[as], asph);
      tc_fence_after();
      epilogue_tile<NT, MB, NEPI, EM, ActT>(a, tile, nt, tmem_base + (uint32_t)(as * ACC_COLS), stage, warp, lane);
      tc_fence_before();
      mbar_arrive_warp(&acc_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

template <int NT, int MB, int ST, int NEPI, int EM, typename ActT>
int launch_gemm_em(const ConvArgs& a, const ConvLayer& L, uint32_t idesc, int grid, cudaStream_t st) {
  constexpr int smem = ST * (MB * 128 * 128 + NT * 128) + (2 * ST + 4) * 8 + 16 + NEPI * 32 * kStageLd * 4;
  static_assert(smem <= 232448, "shared memory budget exceeded");
  static_assert(NEPI % 4 == 0 && (MB * (NT / 32)) % (NEPI / 4) == 0, "epilogue warps must divide the column blocks");
  static bool configured = false;
  if (!configured) {
    VT_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<NT, MB, ST, NEPI, EM, ActT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  VT_CUDA_OK(launch_pdl(k_gemm_tc<NT, MB, ST, NEPI, EM, ActT>, dim3((unsigned)grid), dim3((NEPI + 2 + kProdWarps) * 32), (size_t)smem, st,
                        a, reinterpret_cast<const uint8_t*>(L.w_gemm), reinterpret_cast<const KBlock*>(L.d_kb), L.n_kb, idesc));
  VT_LAUNCHED();
  return VT_OK;
}

template <int NT, typename ActT>
int launch_gemm_nt(const ConvArgs& a, const ConvLayer& L, uint32_t idesc, int grid, cudaStream_t st) {
  const bool split = a.act[0].dst && a.act[1].dst && a.act[0].kind == ACT_ELU && a.act[1].kind == ACT_ELU;
  if (split && !a.out) return launch_gemm_em<NT, 2, 4, 4, EM_ELU | EM_SPLIT, ActT>(a, L, idesc, grid, st);
  if (a.out && a.act[0].kind == ACT_ELU && !a.act[0].dst)
    return launch_gemm_em<NT, 2, 4, 4, EM_ELU | EM_OUT, ActT>(a, L, idesc, grid, st);
  if (a.out && a.act[0].dst && a.act[0].kind == ACT_LRELU && a.act_from_out && !a.act[1].dst)
    return launch_gemm_em<NT, 2, 4, 4, EM_OUT | EM_OACT, ActT>(a, L, idesc, grid, st);
  // layers with few K blocks (the source_downs) are bound by their epilogue: 8 epilogue warps, 3 stages
  const bool epi_heavy = L.n_kb <= 12;
  if (a.out && a.act[0].dst && a.act[0].kind == ACT_SNAKE && !a.act[1].dst)
    return epi_heavy ? launch_gemm_em<NT, 2, 3, 8, EM_OUT | EM_ACT1, ActT>(a, L, idesc, grid, st)
                     : launch_gemm_em<NT, 2, 4, 4, EM_OUT | EM_ACT1, ActT>(a, L, idesc, grid, st);
  if (a.out && !a.act[0].dst && a.act[0].kind == ACT_NONE && !a.act[1].dst)
    return epi_heavy ? launch_gemm_em<NT, 2, 3, 8, EM_OUT, ActT>(a, L, idesc, grid, st)
                     : launch_gemm_em<NT, 2, 4, 4, EM_OUT, ActT>(a, L, idesc, grid, st);
  VT_REQUIRE(false, "gemm_tc: no compiled epilogue for layer %s", L.name.c_str());
  return VT_OK;
}

}  // namespace tc

// Pack the K-block weight matrices ([n_kb][64][N] fp32; part 0 = rounded value, 1 = rounding residual) into
// the per (column tile, K block) shared-memory images [NT rows of 128 B, 16-byte chunks XOR-swizzled].
int pack_gemm_tc(ConvLayer& L, const std::vector<KBlock>& kbs, const std::vector<int>& part, const std::vector<float>& wkb,
                 int NT, int elem, std::vector<void*>& allocs) {
  const int N = L.cout, n_kb = (int)kbs.size();
  VT_REQUIRE(N % NT == 0 && (NT == 64 || NT == 128) && n_kb > 0 && wkb.size() == (size_t)n_kb * 64 * N && (int)part.size() == n_kb,
             "gemm_tc: layer %s does not tile (cout=%d)", L.name.c_str(), N);
  VT_REQUIRE(elem == ELEM_F16 || elem == ELEM_BF16, "gemm_tc: operand type must be fp16 or bf16");
  auto round_f = [&](float v) {
    return elem == ELEM_F16 ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16_rn(v));
  };
  auto to_bits = [&](float v) {
    uint16_t bits;
    if (elem == ELEM_F16) { const __half hv = __float2half_rn(v); std::memcpy(&bits, &hv, 2); }
    else { const __nv_bfloat16 bv = __float2bfloat16_rn(v); std::memcpy(&bits, &bv, 2); }
    return bits;
  };
  std::vector<uint16_t> img((size_t)n_kb * 64 * N);
  size_t chunk = 0;
  for (int nt = 0; nt < N / NT; ++nt)
    for (int kb = 0; kb < n_kb; ++kb, ++chunk) {
      uint16_t* dst = img.data() + chunk * (size_t)NT * 64;
      for (int co = 0; co < NT; ++co)
        for (int e = 0; e < 64; ++e) {
          float v = wkb[((size_t)kb * 64 + e) * N + nt * NT + co];
          if (part[kb]) v = v - round_f(v);
          dst[(size_t)co * 64 + (size_t)((((e >> 3) ^ (co & 7)) << 3) + (e & 7))] = to_bits(v);
        }
    }
  void* p = nullptr;
  VT_CUDA_OK(cudaMalloc(&p, img.size() * 2));
  allocs.push_back(p);
  VT_CUDA_OK(cudaMemcpy(p, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  void* q = nullptr;
  VT_CUDA_OK(cudaMalloc(&q, kbs.size() * sizeof(KBlock)));
  allocs.push_back(q);
  VT_CUDA_OK(cudaMemcpy(q, kbs.data(), kbs.size() * sizeof(KBlock), cudaMemcpyHostToDevice));
  L.w_gemm = p;
  L.d_kb = q;
  L.n_kb = n_kb;
  L.gemm_nt = NT;
  L.gemm_elem = elem;
  return VT_OK;
}

// `a.tiles` must hold tiles of at most 256 output steps.
int launch_gemm_tc(const ConvArgs& a, const ConvLayer& L, int act_elem, cudaStream_t st) {
  VT_REQUIRE(L.w_gemm && a.in_act && a.tiles, "gemm_tc: layer %s not packed", L.name.c_str());
  if (a.n_tiles == 0) return VT_OK;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    VT_CUDA_OK(cudaGetDevice(&dev));
    VT_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int NT = L.gemm_nt;
  const long long total = (long long)a.n_tiles * (L.cout / NT);
  const int grid = total < sm_count ? (int)total : sm_count;
  const uint32_t fmt = L.gemm_elem == ELEM_F16 ? 0u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
  // ActT is the type of the activation copies the epilogue writes (the vocoder's operand type); the MMA
  // operand format comes from the layer (split-precision layers are always fp16)
  if (act_elem == ELEM_F16)
    return NT == 64 ? tc::launch_gemm_nt<64, __half>(a, L, idesc, grid, st) : tc::launch_gemm_nt<128, __half>(a, L, idesc, grid, st);
  return NT == 64 ? tc::launch_gemm_nt<64, __nv_bfloat16>(a, L, idesc, grid, st)
                  : tc::launch_gemm_nt<128, __nv_bfloat16>(a, L, idesc, grid, st);
}

}  // namespace vt
