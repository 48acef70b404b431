// Tensor-core convolution for the ResBlock layers (95 % of the path's FLOPs): implicit GEMM on
// tcgen05 with TMEM accumulators, sm_100a only.
//
//   out[q][co] = sum_j sum_ci act[q + j*dil - pad][ci] * W[co][ci][j]        C_in = C_out = C in {64,128,256}
//
// GEMM view per CTA tile: M = 128*MB time steps, N = C output channels, K = C*k.
//   A (activations, fp16/bf16, channel-last rows) : the tile *with its halo* (128*MB + (k-1)*dil rows) is
//      brought into shared memory ONCE as K-major no-swizzle panels [C/8][RA rows][8 ch]; every tap is the
//      same panels read through a shared-memory descriptor whose start address is shifted by j*dil rows
//      (16 B per row), so a k-tap conv costs one activation load, not k.  RA is odd => the 16-byte cp.async
//      scatter that builds the panels is bank-conflict free.
//   B (weights) : pre-packed on the host per (tap, 64-channel block) as the exact shared-memory image
//      [8][C][8 ch]; streamed with 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx) through a ring.
//   D (fp32) : TMEM, double buffered (2 x MB x C columns) so the epilogue of tile i overlaps the MMAs of i+1.
//
// Warp roles (256 threads, persistent over a static round-robin tile schedule):
//   warps 0-3 epilogue (tcgen05.ld -> bias/residual/Snake -> global), warp 4 MMA issuer + TMEM owner,
//   warp 5 weight producer (bulk TMA), warps 6-7 activation producers (cp.async).
#include "vt_tc.cuh"

#include <cstdlib>
#include <cstring>

namespace vt {

namespace tc {

// CIN input channels (GEMM K per tap), NT output columns per CTA tile (GEMM N), MB 128-row M blocks
// per tile, HALO = largest (k-1)*dil the instance accepts.
template <int CIN, int NT, int MB, int HALO>
struct Cfg {
  // rows of the activation tile: a multiple of 8 (1024-byte swizzle atoms) or odd (no-swizzle panels)
  static constexpr int RA = kSwz ? 128 * MB + ((HALO + 7) / 8) * 8 : 128 * MB + HALO + 1;
  static constexpr int KC = CIN / 8;                // 16-byte K chunks per row
  static constexpr int A_BYTES = KC * RA * 16;      // = (CIN/64) blocks x RA rows x 128 B when swizzled
  static constexpr int W_BYTES = NT * 128;          // one (n-tile, tap, 64-channel block) weight chunk
  static constexpr int ACC_COLS = MB * NT;
  static constexpr int TMEM_COLS = pow2_cols(2 * ACC_COLS);
  static_assert(HALO % 2 == 0 && CIN % 64 == 0 && NT % 32 == 0 && 2 * ACC_COLS <= 512, "bad tile configuration");
};


template <int CIN, int NT, int MB, int A_ST, int W_ST, int NEPI, int HALO, int EM, typename ActT>
__global__ void __launch_bounds__((NEPI + 2 + kProdWarps) * 32, 1)
k_conv_tc(const ConvArgs a, const uint8_t* __restrict__ wtc, const uint32_t idesc) {
  using K = Cfg<CIN, NT, MB, HALO>;
  constexpr int W_MMA = NEPI, W_WP = NEPI + 1, W_AP = NEPI + 2;   // warp roles after the epilogue warps
  constexpr int CB = CIN / 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = sA + A_ST * K::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + W_ST * K::W_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + A_ST * CB;        // one (full, empty) pair per 64-channel block of an A stage
  uint64_t* w_full = a_empty + A_ST * CB;
  uint64_t* w_empty = w_full + W_ST;
  uint64_t* acc_full = w_empty + W_ST;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* stage_all = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tmem_slot) + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < A_ST * CB; ++i) { mbar_init(&a_full[i], kProdWarps); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < W_ST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NEPI); }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)K::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();                             // dependents may be scheduled (they take an SM when its CTA of this grid exits)
  if (warp != W_WP) pdl_wait();                // everything but the (static) weight stream waits for the previous kernel

  const int nchunks = a.k * CB;                 // weight chunks per tile
  const int n_nt = a.cout / NT;                 // column tiles
  const int n_tiles = a.n_tiles * n_nt;         // (row tile, column tile) pairs, column tile fastest

  if (warp >= W_AP) {
    // ---------------- activation producers: halo tile -> K-major panels
    const int pt = threadIdx.x - W_AP * 32;
    const ActT* in = reinterpret_cast<const ActT*>(a.in_act);
    const int r_need = 128 * MB + (a.k - 1) * a.dil;
    const int pieces = r_need * K::KC;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int ab = it % A_ST;
      const uint32_t ph = (uint32_t)(it / A_ST) & 1u;
      const ConvTile tile = a.tiles[t / n_nt];
      const ActT* src = in + (tile.in_row0 + tile.q0 - a.pad) * (long long)CIN;
      const uint32_t dst = smem_u32(sA + ab * K::A_BYTES);
      // one 64-channel block at a time: the MMA warp walks the blocks in the same order (block outer, taps
      // inner) and releases each block as soon as its taps are issued, so the refill of block 0 for the next
      // tile overlaps the MMAs of this tile's later blocks even with a single A stage
      constexpr int KCB = (kSwz && A_ST == 1) ? 8 : K::KC;      // 16-byte chunks per row of a block
      constexpr int NB = (kSwz && A_ST == 1) ? CB : 1;          // with two A stages whole tiles already overlap
      const int bpieces = r_need * KCB;
      for (int cb = 0; cb < NB; ++cb) {
        mbar_wait(&a_empty[ab * CB + cb], ph ^ 1u);
        if (NB == 1) for (int x = 1; x < CB; ++x) mbar_wait(&a_empty[ab * CB + x], ph ^ 1u);
        if (pt == 0 && cb == 0) trace_ev(a.trace, it, 0);
        for (int p = pt; p < bpieces && !(a.dbg & 2); p += kProd) {
          const int r = p / KCB, c = cb * KCB + (p - r * KCB);
          const uint32_t off = kSwz ? (uint32_t)((c >> 3) * K::RA + r) * 128u + (uint32_t)(((c & 7) ^ (r & 7)) << 4)
                                    : (uint32_t)(c * K::RA + r) * 16u;
          cp_async16(dst + off, src + (long long)r * CIN + c * 8);
        }
        if (pt == 0 && cb == NB - 1) trace_ev(a.trace, it, 1);
        cp_async_wait_all();
        fence_proxy_async();
        if (pt == 0 && cb == NB - 1) trace_ev(a.trace, it, 2);
        mbar_arrive_warp(&a_full[ab * CB + cb]);
        if (NB == 1) for (int x = 1; x < CB; ++x) mbar_arrive_warp(&a_full[ab * CB + x]);
      }
    }
  } else if (warp == W_WP) {
    // ---------------- weight producer: one bulk copy per (tap, 64-channel block) of this tile's column tile
    if (lane == 0) {
      uint32_t ws = 0, ph = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int nt = t % n_nt;
        const uint8_t* wsrc = wtc + (size_t)nt * nchunks * K::W_BYTES;
        // issue order of the MMA warp: channel block outer, taps inner; the packed image is (tap, block)
        for (int q = 0; q < nchunks; ++q) {
          const int cb = q / a.k, j = q - cb * a.k;
          const int c = j * CB + cb;
          if ((a.tap_skip >> (4 * nt + j)) & 1ull) continue;   // all-zero tap of this column tile: not loaded, not issued
          mbar_wait(&w_empty[ws], ph ^ 1u);
          if (a.dbg & 1) mbar_arrive(&w_full[ws]);
          else {
            mbar_arrive_expect_tx(&w_full[ws], K::W_BYTES);
            bulk_g2s(sW + ws * K::W_BYTES, wsrc + (size_t)c * K::W_BYTES, K::W_BYTES, &w_full[ws]);
          }
          if (++ws == (uint32_t)W_ST) { ws = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ---------------- MMA issuer: the whole warp runs the loop (convergent waits), one elected lane
    // issues.  Descriptors are (lo, hi) pairs: hi is a constant of the layout, lo advances by adds.
    static_assert(kSwz, "the issue loop assumes the SWIZZLE_128B operand layout");
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = ((smem_u32(sA) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t w_lo0 = ((smem_u32(sW) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t tap16 = (uint32_t)a.dil * 8u;                 // one tap = dil rows of 128 B, in 16-byte units
    const bool mma_on = !(a.dbg & 16);
    uint32_t ws = 0, wph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int ab = it % A_ST;
      const uint32_t aph = (uint32_t)(it / A_ST) & 1u;
      const int as = it & 1;
      const uint32_t asph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&acc_empty[as], asph ^ 1u);
      if (lane == 0) trace_ev(a.trace, it, 3);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(as * K::ACC_COLS);
      const uint32_t a_tile = a_lo0 + (uint32_t)ab * (uint32_t)(K::A_BYTES >> 4);
      uint32_t acc = 0;
      long long w_wait = 0;
#pragma unroll 1
      for (int cb = 0; cb < CB; ++cb) {
        mbar_wait(&a_full[ab * CB + cb], aph);
        if (lane == 0 && cb == 0) trace_ev(a.trace, it, 4);
        tc_fence_after();
        uint32_t a_chunk = a_tile + (uint32_t)cb * (uint32_t)K::RA * 8u;
        for (int j = 0; j < a.k; ++j, a_chunk += tap16) {
          if ((a.tap_skip >> (4 * (t % n_nt) + j)) & 1ull) continue;
          if (a.trace) {
            const long long tw = clock64();
            mbar_wait(&w_full[ws], wph);
            w_wait += clock64() - tw;
          } else {
            mbar_wait(&w_full[ws], wph);
          }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_lo = w_lo0 + ws * (uint32_t)(K::W_BYTES >> 4);
            if (mma_on) {
#pragma unroll
              for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_f16_lh(d0 + (uint32_t)(mb * NT), a_chunk + (uint32_t)(mb * 1024 + ks * 2), b_lo + (uint32_t)(ks * 2),
                              kDescHi, idesc, ks == 0 ? acc : 1u);
            }
            umma_commit(&w_empty[ws]);
          }
          __syncwarp();
          acc = 1u;
          if (++ws == (uint32_t)W_ST) { ws = 0; wph ^= 1u; }
        }
        if (elect_one()) umma_commit(&a_empty[ab * CB + cb]);     // this block's taps are issued: it may be refilled
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_full[as]);
      __syncwarp();
      if (lane == 0) trace_ev(a.trace, it, 5);
      if (lane == 0 && a.trace && blockIdx.x == 0 && it < kTraceTiles) a.trace[it * kTraceEvents + 8] = w_wait;
    }
  } else {
    // ---------------- epilogue warps 0..NEPI-1.  Warp w may touch TMEM lanes 32*(w%4)..+31; the NEPI/4
    // warps of a lane quarter split the tile's 32-column blocks round-robin.
    // tcgen05.ld gives every thread 32 columns of ITS row; global memory wants a warp to touch whole
    // rows.  Each 32x32 fp32 block is transposed through a padded shared-memory stage so that every
    // global instruction covers 4 rows x 128 contiguous bytes (4 wavefronts instead of 32).
    float* stage = stage_all + warp * (32 * kStageLd);
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t asph = (uint32_t)(it >> 1) & 1u;
      const int rt = t / n_nt, nt = t - rt * n_nt;
      const ConvTile tile = a.tiles[rt];
      mbar_wait(&acc_full[as], asph);
      if (threadIdx.x == 0) trace_ev(a.trace, it, 6);
      tc_fence_after();
      epilogue_tile<NT, MB, NEPI, EM, ActT>(a, tile, nt, tmem_base + (uint32_t)(as * K::ACC_COLS), stage, warp, lane);
      tc_fence_before();
      mbar_arrive_warp(&acc_empty[as]);
      if (threadIdx.x == 0) trace_ev(a.trace, it, 7);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)K::TMEM_COLS) : "memory");
  }
}

template <int CIN, int NT, int MB, int A_ST, int W_ST, int NEPI, int HALO>
constexpr int smem_bytes() {
  return A_ST * Cfg<CIN, NT, MB, HALO>::A_BYTES + W_ST * Cfg<CIN, NT, MB, HALO>::W_BYTES +
         (2 * A_ST * (CIN / 64) + 2 * W_ST + 4) * 8 + 16 + NEPI * 32 * kStageLd * 4;
}

template <int CIN, int NT, int MB, int A_ST, int W_ST, int NEPI, int HALO, int EM, typename ActT>
int launch_em(const ConvArgs& a, const void* wtc, uint32_t idesc, int grid, cudaStream_t st) {
  constexpr int smem = smem_bytes<CIN, NT, MB, A_ST, W_ST, NEPI, HALO>();
  static_assert(smem <= 232448, "shared memory budget exceeded");
  static_assert(NEPI % 4 == 0 && (MB * (NT / 32)) % (NEPI / 4) == 0, "epilogue warps must divide the column blocks");
  static bool configured = false;
  if (!configured) {
    VT_CUDA_OK(cudaFuncSetAttribute(k_conv_tc<CIN, NT, MB, A_ST, W_ST, NEPI, HALO, EM, ActT>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    VT_CUDA_OK(cudaFuncSetAttribute(k_conv_tc<CIN, NT, MB, A_ST, W_ST, NEPI, HALO, EM, ActT>,
                                    cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  VT_CUDA_OK(launch_pdl(k_conv_tc<CIN, NT, MB, A_ST, W_ST, NEPI, HALO, EM, ActT>, dim3((unsigned)grid), dim3((NEPI + 2 + kProdWarps) * 32),
                        (size_t)smem, st, a, reinterpret_cast<const uint8_t*>(wtc), idesc));
  VT_LAUNCHED();
  return VT_OK;
}

// ResBlock convolutions: map the runtime epilogue description onto a compiled specialisation.
template <int C, int MB, int A_ST, int W_ST, int NEPI, typename ActT>
int launch_resblock(const ConvArgs& a, const void* wtc, uint32_t idesc, int grid, cudaStream_t st) {
  int nsnake = 0;
  while (nsnake < 3 && a.act[nsnake].dst && a.act[nsnake].kind == ACT_SNAKE) ++nsnake;
  const bool oact = nsnake == 0 && a.act[0].dst && a.act[0].kind == ACT_LRELU && a.act_from_out;
  for (int s = nsnake + (oact ? 1 : 0); s < 3; ++s)
    VT_REQUIRE(!a.act[s].dst, "conv_tc: unsupported activation-copy combination");
  const bool r1 = a.res1 != nullptr, r2 = a.res2 != nullptr, out = a.out != nullptr;
  if (!r1 && !r2 && !out && nsnake == 1)
    return launch_em<C, C, MB, A_ST, W_ST, NEPI, 50, EM_ACT1, ActT>(a, wtc, idesc, grid, st);
  if (r1 && !r2 && out && !a.out_accum && nsnake == 1)
    return launch_em<C, C, MB, A_ST, W_ST, NEPI, 50, EM_RES1 | EM_OUT | EM_ACT1, ActT>(a, wtc, idesc, grid, st);
  if (r1 && r2 && out && !a.out_accum && nsnake == 3)
    return launch_em<C, C, MB, A_ST, W_ST, NEPI, 50, EM_RES1 | EM_RES2 | EM_OUT | EM_ACT3, ActT>(a, wtc, idesc, grid, st);
  if (r1 && !r2 && out && nsnake == 0 && !oact)
    return launch_em<C, C, MB, A_ST, W_ST, NEPI, 50, EM_RES1 | EM_OUT | EM_ACCUM, ActT>(a, wtc, idesc, grid, st);
  if (r1 && !r2 && out && nsnake == 0 && oact)
    return launch_em<C, C, MB, A_ST, W_ST, NEPI, 50, EM_RES1 | EM_OUT | EM_ACCUM | EM_OACT, ActT>(a, wtc, idesc, grid, st);
  VT_REQUIRE(false, "conv_tc: no compiled epilogue for res1=%d res2=%d out=%d accum=%d snake=%d", (int)r1, (int)r2, (int)out,
             a.out_accum, nsnake);
  return VT_OK;
}

// Plain layers (transposed convs as 3-tap convs, conv_pre, conv_post): fp32 output, optional leaky-ReLU copy.
template <int CIN, int NT, int MB, int A_ST, int W_ST, int NEPI, int HALO, typename ActT>
int launch_plain(const ConvArgs& a, const void* wtc, uint32_t idesc, int grid, cudaStream_t st) {
  VT_REQUIRE(a.out && !a.res1 && !a.res2 && !a.out_accum && !a.act[1].dst && !a.act[2].dst,
             "conv_tc: plain layers write one fp32 output");
  if (a.act[0].dst) {
    VT_REQUIRE(a.act[0].kind == ACT_LRELU && a.act_from_out, "conv_tc: plain layers support a leaky-ReLU output copy only");
    return launch_em<CIN, NT, MB, A_ST, W_ST, NEPI, HALO, EM_OUT | EM_OACT, ActT>(a, wtc, idesc, grid, st);
  }
  return launch_em<CIN, NT, MB, A_ST, W_ST, NEPI, HALO, EM_OUT, ActT>(a, wtc, idesc, grid, st);
}

}  // namespace tc

// Which compiled instance a layer maps to (0 = none): 1..3 ResBlock C=64/128/256, 4..6 the three
// transposed convs, 7 conv_post (C = 64), 8 conv_pre, 9 conv_post at C = 128 (two-stage generators: CosyVoice-300M).
static int tc_instance(const ConvLayer& L) {
  if (L.stride != 1 || L.pad > kGap) return 0;
  if (L.out_mul == 1 && L.cin == L.cout && (L.k - 1) * L.dil <= 50) {
    if (L.cin == 64 && L.k > 1) return 1;
    if (L.cin == 128) return 2;
    if (L.cin == 256) return 3;
  }
  if (L.out_mul > 1 && L.k == 3 && L.dil == 1) {
    if (L.cin == 512 && L.phase_c == 256) return 4;
    if (L.cin == 256 && L.phase_c == 128) return 5;
    if (L.cin == 128 && L.phase_c == 64) return 6;
  }
  if (L.out_mul == 1 && L.cin == 64 && L.cout == 32 && (L.k - 1) * L.dil <= 6) return 7;
  if (L.out_mul == 1 && L.cin == 128 && L.cout == 512 && (L.k - 1) * L.dil <= 6) return 8;
  if (L.out_mul == 1 && L.cin == 128 && L.cout == 32 && (L.k - 1) * L.dil <= 6) return 9;   // conv_post of a two-stage generator
  return 0;
}

static int tc_nt(int inst) {
  switch (inst) {
    case 1: return 64;
    case 2: return 128;
    case 3: return 256;
    case 4: return 128;
    case 5: return 128;
    case 6: return 64;
    case 7: return 32;
    case 8: return 128;
    case 9: return 32;
  }
  return 0;
}

bool conv_tc_supported(const ConvLayer& L) { return tc_instance(L) != 0; }

int conv_tc_tile_rows(const ConvLayer& L) {
  const int inst = tc_instance(L);
  return (inst == 1 || inst == 2 || inst == 7 || inst == 9) ? 256 : 128;
}

// [k][cin][cout] fp32 -> per (column tile nt, tap j, 64-channel block cb) the shared-memory image
// [8][NT][8] in the operand type
int pack_conv_tc(ConvLayer& L, const std::vector<float>& w, int act_elem, std::vector<void*>& allocs) {
  const int inst = tc_instance(L);
  VT_REQUIRE(inst != 0, "conv_tc: layer %s has no tensor-core instance", L.name.c_str());
  const int CIN = L.cin, N = L.cout, k = L.k, NT = tc_nt(inst);
  VT_REQUIRE(N % NT == 0 && CIN % 64 == 0, "conv_tc: layer %s does not tile (cin=%d cout=%d)", L.name.c_str(), CIN, N);
  const size_t n = (size_t)k * CIN * N;
  std::vector<uint16_t> img(n);
  auto to_bits = [&](float v) {
    uint16_t bits;
    if (act_elem == ELEM_F16) {
      const __half hv = __float2half_rn(v);
      std::memcpy(&bits, &hv, 2);
    } else {
      const __nv_bfloat16 bv = __float2bfloat16_rn(v);
      std::memcpy(&bits, &bv, 2);
    }
    return bits;
  };
  size_t chunk = 0;
  for (int nt = 0; nt < N / NT; ++nt)
    for (int j = 0; j < k; ++j)
      for (int cb = 0; cb < CIN / 64; ++cb, ++chunk) {
        uint16_t* dst = img.data() + chunk * (size_t)NT * 64;
        for (int co = 0; co < NT; ++co)
          for (int k8 = 0; k8 < 8; ++k8)
            for (int e = 0; e < 8; ++e) {
              const float v = w[((size_t)j * CIN + cb * 64 + k8 * 8 + e) * N + nt * NT + co];
              // swizzled: row co is 128 B, its 16-byte chunk k8 sits at k8 ^ (co & 7); else panels [k8][co][e]
              const size_t pos = tc::kSwz ? (size_t)co * 64 + (size_t)((k8 ^ (co & 7)) * 8 + e)
                                          : ((size_t)k8 * NT + co) * 8 + e;
              dst[pos] = to_bits(v);
            }
      }
  // phase-decomposed transposed convs: a tap whose weights are all zero for a whole column tile is skipped
  L.tap_skip = 0;
  if (L.out_mul > 1 && N / NT <= 16 && k <= 4) {
    for (int nt = 0; nt < N / NT; ++nt)
      for (int j = 0; j < k; ++j) {
        bool zero = true;
        for (int ci = 0; ci < CIN && zero; ++ci)
          for (int co = 0; co < NT; ++co)
            if (w[((size_t)j * CIN + ci) * N + nt * NT + co] != 0.0f) { zero = false; break; }
        if (zero) L.tap_skip |= 1ull << (4 * nt + j);
      }
    for (int nt = 0; nt < N / NT; ++nt)
      VT_REQUIRE(((L.tap_skip >> (4 * nt)) & 15ull) != ((1ull << k) - 1), "conv_tc: layer %s has an all-zero column tile", L.name.c_str());
  }
  void* p = nullptr;
  VT_CUDA_OK(cudaMalloc(&p, n * 2));
  allocs.push_back(p);
  VT_CUDA_OK(cudaMemcpy(p, img.data(), n * 2, cudaMemcpyHostToDevice));
  L.w_tc = p;
  return VT_OK;
}

template <typename ActT>
static int launch_conv_tc_t(const ConvArgs& a, const ConvLayer& L, int inst, uint32_t idesc, int grid, cudaStream_t st) {
  switch (inst) {
    case 1: return tc::launch_resblock<64, 2, 2, 4, 16, ActT>(a, L.w_tc, idesc, grid, st);
    case 2: return tc::launch_resblock<128, 2, 1, 4, 16, ActT>(a, L.w_tc, idesc, grid, st);
    case 3: return tc::launch_resblock<256, 1, 1, 3, 8, ActT>(a, L.w_tc, idesc, grid, st);
    case 4: return tc::launch_plain<512, 128, 1, 1, 4, 4, 2, ActT>(a, L.w_tc, idesc, grid, st);
    case 5: return tc::launch_plain<256, 128, 1, 2, 4, 4, 2, ActT>(a, L.w_tc, idesc, grid, st);
    case 6: return tc::launch_plain<128, 64, 1, 2, 4, 8, 2, ActT>(a, L.w_tc, idesc, grid, st);
    case 7: return tc::launch_plain<64, 32, 2, 2, 4, 8, 6, ActT>(a, L.w_tc, idesc, grid, st);
    case 16: return tc::launch_plain<128, 64, 1, 2, 3, 4, 2, ActT>(a, L.w_tc, idesc, grid, st);  // 110 KB: two CTAs per SM
    case 17: return tc::launch_plain<64, 32, 2, 2, 4, 4, 6, ActT>(a, L.w_tc, idesc, grid, st);   // 102 KB: two CTAs per SM
    case 8: return tc::launch_plain<128, 128, 1, 2, 4, 4, 6, ActT>(a, L.w_tc, idesc, grid, st);
    case 9: return tc::launch_plain<128, 32, 2, 2, 4, 4, 6, ActT>(a, L.w_tc, idesc, grid, st);   // CosyVoice-300M conv_post (128 -> 18)
  }
  VT_REQUIRE(false, "conv_tc: layer %s has no tensor-core instance", L.name.c_str());
  return VT_OK;
}

int launch_conv_tc(const ConvArgs& a_in, const ConvLayer& L, int act_elem, const void* tc_tiles, int n_tc_tiles,
                   int tile_rows, cudaStream_t st) {
  VT_REQUIRE(L.w_tc && a_in.in_act && (act_elem == ELEM_F16 || act_elem == ELEM_BF16), "conv_tc: layer %s not packed", L.name.c_str());
  if (n_tc_tiles == 0) return VT_OK;
  const int inst = tc_instance(L);
  const int NT = tc_nt(inst);
  VT_REQUIRE(inst != 0 && tile_rows == conv_tc_tile_rows(L), "conv_tc: tile table does not match layer %s", L.name.c_str());
  ConvArgs a = a_in;
  a.tap_skip = L.tap_skip;
  a.tiles = reinterpret_cast<const ConvTile*>(tc_tiles);
  a.n_tiles = n_tc_tiles;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    VT_CUDA_OK(cudaGetDevice(&dev));
    VT_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  // debug timeline: VT_TC_TRACE=<layer name> dumps CTA 0's per-tile role timestamps to stderr
  static const char* trace_name = getenv("VT_TC_TRACE");
  static long long* d_trace = nullptr;
  const bool tracing = trace_name && L.name == trace_name;
  // timing ablation (results are wrong): VT_TC_DBG bits 1 = no weight copies, 2 = no activation copies,
  // 4 = no epilogue, 8 = epilogue without global traffic, 16 = no MMAs
  static const int dbg = getenv("VT_TC_DBG") ? atoi(getenv("VT_TC_DBG")) : 0;
  a.dbg = dbg;
  if (tracing) {
    if (!d_trace) VT_CUDA_OK(cudaMalloc(&d_trace, tc::kTraceTiles * tc::kTraceEvents * 8));
    VT_CUDA_OK(cudaMemsetAsync(d_trace, 0, tc::kTraceTiles * tc::kTraceEvents * 8, st));
    a.trace = d_trace;
  }
  const long long total = (long long)n_tc_tiles * (L.cout / NT);
  // conv_post (64 -> 18 channels, instance 7) and the last transposed conv (128 -> 3 x 64, instance 6) are
  // latency-bound pipelines at ~30 % issue and 30-35 % DRAM utilisation with one CTA per SM.  With 4 instead of 8
  // epilogue warps (and a 3-deep weight ring for instance 6) a CTA needs 102 / 110 KB of shared memory and 128 TMEM
  // columns, so two CTAs share an SM and hide each other's stalls: conv_post 0.32 -> 0.21 ms.
  // VT_CONV_2CTA=0: one CTA per SM.
  static const bool two_cta = !(getenv("VT_CONV_2CTA") && getenv("VT_CONV_2CTA")[0] == '0');
  const int launch_inst = ((inst == 7 || inst == 6) && two_cta) ? inst + 10 : inst;
  const int slots = launch_inst >= 16 ? 2 * sm_count : sm_count;
  const int grid = total < slots ? (int)total : slots;
  const uint32_t fmt = act_elem == ELEM_F16 ? 0u : 1u;
  // instruction descriptor: D = f32 (bits 4-5 = 1), A/B format (bits 7-9, 10-12), K-major A and B,
  // N>>3 at bits 17-22, M>>4 at bits 24-28
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
  const int rc = act_elem == ELEM_F16 ? launch_conv_tc_t<__half>(a, L, launch_inst, idesc, grid, st)
                                      : launch_conv_tc_t<__nv_bfloat16>(a, L, launch_inst, idesc, grid, st);
  if (tracing && rc == VT_OK) {
    std::vector<long long> h(tc::kTraceTiles * tc::kTraceEvents);
    VT_CUDA_OK(cudaStreamSynchronize(st));
    VT_CUDA_OK(cudaMemcpy(h.data(), d_trace, h.size() * 8, cudaMemcpyDeviceToHost));
    long long t0 = 0;
    for (size_t q = 0; q < h.size(); ++q)
      if ((int)(q % tc::kTraceEvents) < 8 && h[q] && (!t0 || h[q] < t0)) t0 = h[q];
    fprintf(stderr, "[vt trace] layer %s k=%d dil=%d tiles=%lld grid=%d  (cycles since first event; "
            "A:wait_empty issued landed | MMA:acc_free a_full issued | EPI:acc_full done)\n", L.name.c_str(), L.k, L.dil, total, grid);
    for (int it = 0; it < tc::kTraceTiles; ++it) {
      if (!h[it * tc::kTraceEvents + 6]) break;
      fprintf(stderr, "[vt trace] %2d", it);
      for (int e = 0; e < 8; ++e) fprintf(stderr, " %8lld", h[it * tc::kTraceEvents + e] ? h[it * tc::kTraceEvents + e] - t0 : -1);
      fprintf(stderr, "  w_wait=%lld", h[it * tc::kTraceEvents + 8]);
      fprintf(stderr, "\n");
    }
  }
  return rc;
}

}  // namespace vt
