// Transposed tcgen05 convolution for the C = 256 ResBlock layers:  D^T[co][t] = sum_j W_j^T[co][ci] X[t + j*dil][ci].
//
// The weight slab is the M = 128 operand (one half of the 256 output channels at a time), the 256 time steps of
// the tile the N operand.  Against the activation-resident kernel of vt_conv_tc.cu (M = 128 time steps, N = 256
// channels) the operand bytes per MMA are the same, but
//   * a weight chunk is streamed from L2 once per 256 time steps instead of once per 128 (weight streaming is the
//     largest single cost of the forward), in 16 KB pieces through a 4-deep ring;
//   * accumulators come out as lane = channel, column = time step: a warp's lanes are 32 consecutive channels of
//     ONE time step, so every global access of the epilogue (residual reads, fp32 stream, fp16 operand copies) is
//     a coalesced line with no shared-memory transposition and no staging buffer;
//   * the two channel halves are the two TMEM accumulator buffers: the epilogue of half 0 runs under the MMAs of
//     half 1, the epilogue of half 1 under the MMAs of the next tile's half 0.
// The activation tile (256 + halo rows x 256 channels fp16 = 156 KB, SWIZZLE_128B) is loaded once per tile in four
// 64-channel blocks with their own barriers; the MMA warp walks half -> block -> tap and releases a block during
// the second half, so its refill for the next tile overlaps the remaining MMAs.
#include "vt_tc.cuh"

#include <cstdlib>
#include <type_traits>

namespace vt {
namespace tc {

constexpr int kTRows = 256;                 // time steps per tile (MMA N)
constexpr int kTRA = kTRows + 56;           // activation rows in shared memory (halo <= 50, multiple of 8)
constexpr int kTWst = 4;                    // weight ring stages of 16 KB
constexpr int kTEpi = 8;                    // epilogue warps: 4 lane quarters x 2 column halves

template <typename T> __device__ __forceinline__ unsigned short op_bits(float v);
template <> __device__ __forceinline__ unsigned short op_bits<__half>(float v) { return __half_as_ushort(__float2half_rn(v)); }
template <> __device__ __forceinline__ unsigned short op_bits<__nv_bfloat16>(float v) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}

// C = 256: two channel halves per tile (NH = 2), one activation stage.  C = 128: one half, the two TMEM buffers
// alternate between tiles, two activation stages.
template <int C, int EM, typename ActT>
__global__ void __launch_bounds__((kTEpi + 2 + kProdWarps) * 32, 1)
k_convT_tc(const ConvArgs a, const uint8_t* __restrict__ wtc, const uint32_t idesc) {
  constexpr int CB = C / 64, NH = C / 128, NA = C == 128 ? 2 : 1;
  constexpr int A_BLK = kTRA * 128, A_BYTES = CB * A_BLK, W_BYTES = 128 * 128;
  constexpr int W_MMA = kTEpi, W_WP = kTEpi + 1, W_AP = kTEpi + 2;
  constexpr int NACT = (EM & EM_ACT3) ? 3 : ((EM & EM_ACT1) ? 1 : 0);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = sA + NA * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + kTWst * W_BYTES);
  uint64_t* a_full = bars;                       // [stage][block]
  uint64_t* a_empty = a_full + NA * CB;
  uint64_t* w_full = a_empty + NA * CB;
  uint64_t* w_empty = w_full + kTWst;
  uint64_t* acc_full = w_empty + kTWst;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NA * CB; ++i) { mbar_init(&a_full[i], kProdWarps); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kTWst; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kTEpi); }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();                             // dependents may be scheduled (they take an SM when its CTA of this grid exits)
  if (warp != W_WP) pdl_wait();                // everything but the (static) weight stream waits for the previous kernel
  const int n_tiles = a.n_tiles;

  if (warp >= W_AP) {
    // ---------------- activation producers: one 64-channel block (rows x 128 B, swizzled) at a time
    const int pt = threadIdx.x - W_AP * 32;
    const ActT* in = reinterpret_cast<const ActT*>(a.in_act);
    const int r_need = kTRows + (a.k - 1) * a.dil;
    const int bpieces = r_need * 8;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const ConvTile tile = a.tiles[t];
      const ActT* src = in + (tile.in_row0 + tile.q0 - a.pad) * (long long)C;
      const int ab = it % NA;
      const uint32_t aph = (uint32_t)(it / NA) & 1u;
      for (int cb = 0; cb < CB; ++cb) {
        mbar_wait(&a_empty[ab * CB + cb], aph ^ 1u);
        const uint32_t dst = smem_u32(sA + ab * A_BYTES + cb * A_BLK);
        for (int p = pt; p < bpieces && !(a.dbg & 2); p += kProd) {
          const int r = p >> 3, c = p & 7;
          cp_async16(dst + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4), src + (long long)r * C + cb * 64 + c * 8);
        }
        cp_async_wait_all();
        fence_proxy_async();
        mbar_arrive_warp(&a_full[ab * CB + cb]);
      }
    }
  } else if (warp == W_WP) {
    // ---------------- weight producer: order (half, block, tap); the packed image is (tap, block) chunks of 256
    // channel rows x 128 B, the half's 128 rows are a contiguous 16 KB piece of a chunk
    if (lane == 0) {
      uint32_t ws = 0, ph = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
        for (int h = 0; h < NH; ++h)
          for (int cb = 0; cb < CB; ++cb)
            for (int j = 0; j < a.k; ++j) {
              mbar_wait(&w_empty[ws], ph ^ 1u);
              if (a.dbg & 1) mbar_arrive(&w_full[ws]);
              else {
                mbar_arrive_expect_tx(&w_full[ws], W_BYTES);
                bulk_g2s(sW + ws * W_BYTES, wtc + ((size_t)(j * CB + cb) * NH + h) * W_BYTES, W_BYTES, &w_full[ws]);
              }
              if (++ws == (uint32_t)kTWst) { ws = 0; ph ^= 1u; }
            }
    }
  } else if (warp == W_MMA) {
    // ---------------- MMA issuer: A operand = weight piece (M = 128 channels), B operand = 256 activation rows
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = ((smem_u32(sA) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t w_lo0 = ((smem_u32(sW) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t tap16 = (uint32_t)a.dil * 8u;
    const bool mma_on = !(a.dbg & 16);
    uint32_t ws = 0, wph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int ab = it % NA;
      const uint32_t aph = (uint32_t)(it / NA) & 1u;
      for (int h = 0; h < NH; ++h) {
        const int seq = it * NH + h, buf = seq & 1;
        mbar_wait(&acc_empty[buf], ((uint32_t)(seq >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * kTRows);
        uint32_t acc = 0;
#pragma unroll 1
        for (int cb = 0; cb < CB; ++cb) {
          if (h == 0) {
            mbar_wait(&a_full[ab * CB + cb], aph);
            tc_fence_after();
          }
          uint32_t x_lo = a_lo0 + (uint32_t)(ab * (A_BYTES >> 4)) + (uint32_t)cb * (uint32_t)(A_BLK >> 4);
          for (int j = 0; j < a.k; ++j, x_lo += tap16) {
            mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t w_lo = w_lo0 + ws * (uint32_t)(W_BYTES >> 4);
              if (mma_on) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_f16_lh(d0, w_lo + (uint32_t)(ks * 2), x_lo + (uint32_t)(ks * 2), kDescHi, idesc, ks == 0 ? acc : 1u);
              }
              umma_commit(&w_empty[ws]);
            }
            __syncwarp();
            acc = 1u;
            if (++ws == (uint32_t)kTWst) { ws = 0; wph ^= 1u; }
          }
          if (h == NH - 1) {                              // every half has read this block: it may be refilled
            if (elect_one()) umma_commit(&a_empty[ab * CB + cb]);
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit(&acc_full[buf]);
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue: lane = channel, TMEM column = time step
    const int quarter = warp & 3, chalf = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const ConvTile tile = a.tiles[t];
      const long long row_base = (tile.out_row0 + tile.q0) * (long long)C;   // channel-last rows of C channels
      const int last = tile.n - 1;
      for (int h = 0; h < NH; ++h) {
        const int seq = it * NH + h, buf = seq & 1;
        const int co = h * 128 + quarter * 32 + lane;
        const float bias = a.bias[co], wsc = a.wscale[co];
        float al[NACT > 0 ? NACT : 1], ia[NACT > 0 ? NACT : 1];
#pragma unroll
        for (int s = 0; s < NACT; ++s) {
          al[s] = a.act[s].alpha[co];
          ia[s] = __fdividef(1.0f, al[s] + 1e-9f);
        }
        const long long obase = row_base + co;
        constexpr bool kLoads = (EM & (EM_RES1 | EM_RES2 | EM_ACCUM)) != 0;
        const bool accum = ((EM & EM_ACCUM) != 0) && a.out_accum;
        const float inv = 1.0f / a.out_scale;
        // residual terms of the first block: requested before the accumulator wait.
        // kFull (every column of the tile is an output step - all tiles but a sequence's last): no index clamps, no
        // per-element predicates, and every address is one base pointer plus a compile-time offset.  The epilogue of
        // conv2 carries three streams per element (residual in, fp32 out, operand copy out); with clamps and 64-bit
        // index arithmetic per access it took 225 us against 157 us for conv1's MMAs at k = 7.
        float x[32];
        auto issue = [&](int cc, auto full_tag) {
          constexpr bool kFull = decltype(full_tag)::value;
          const int col0 = chalf * 128 + cc * 32;
          const float* rp = a.res1 + obase + (long long)col0 * C;
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            float r = 0.0f;
            if constexpr ((EM & EM_RES1) != 0) {
              if constexpr (kFull) r = __ldg(rp + q * C);
              else r = __ldg(a.res1 + obase + (long long)(col0 + q < last ? col0 + q : last) * C);
            }
            x[q] = r;
          }
        };
        auto block = [&](int cc, auto full_tag) {
          constexpr bool kFull = decltype(full_tag)::value;
          const int col0 = chalf * 128 + cc * 32;
          const long long cbase = obase + (long long)col0 * C;
          if constexpr ((EM & EM_RES2) != 0) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              x[q] += kFull ? __ldg(a.res2 + cbase + q * C) : __ldg(a.res2 + obase + (long long)(col0 + q < last ? col0 + q : last) * C);
          }
          if (accum) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              x[q] = fmaf(kFull ? __ldg(a.out + cbase + q * C) : __ldg(a.out + obase + (long long)(col0 + q < last ? col0 + q : last) * C),
                          inv, x[q]);
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(tmem_base + lane_sel + (uint32_t)(buf * kTRows + col0 + hh * 16))
                : "memory");
            tmem_ld_wait();
            // one output element: everything but the Snake copies (those run two at a time in the full path)
            auto finish = [&](int q, long long idx) {
              float y = fmaf(__uint_as_float(v[q]), wsc, bias);
              if constexpr (kLoads) y += x[hh * 16 + q];
              if constexpr ((EM & EM_OUT) != 0) {
                const float o = y * a.out_scale;
                a.out[idx] = o;
                if constexpr ((EM & EM_OACT) != 0) {
                  const float sl = a.act[0].slope;
                  reinterpret_cast<unsigned short*>(a.act[0].dst)[idx] = op_bits<ActT>(o > 0.f ? o : o * sl);
                }
              }
              return y;
            };
            if constexpr (kFull) {
#pragma unroll
              for (int q = 0; q < 16; q += 2) {
                const long long idx0 = cbase + (hh * 16 + q) * C, idx1 = idx0 + C;
                const float y0 = finish(q, idx0), y1 = finish(q + 1, idx1);
#pragma unroll
                for (int s = 0; s < NACT; ++s) {
                  float z0, z1;
                  snake_f2(y0, y1, al[s], al[s], ia[s], ia[s], z0, z1);
                  reinterpret_cast<unsigned short*>(a.act[s].dst)[idx0] = op_bits<ActT>(z0);
                  reinterpret_cast<unsigned short*>(a.act[s].dst)[idx1] = op_bits<ActT>(z1);
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const int row = col0 + hh * 16 + q;
                if (row < tile.n) {
                  const long long idx = obase + (long long)row * C;
                  const float y = finish(q, idx);
#pragma unroll
                  for (int s = 0; s < NACT; ++s)
                    reinterpret_cast<unsigned short*>(a.act[s].dst)[idx] = op_bits<ActT>(snake_f(y, al[s], ia[s]));
                }
              }
            }
          }
          if constexpr (kLoads) {
            if (cc < 3) issue(cc + 1, full_tag);
          }
        };
        const bool full = tile.n == kTRows && !(a.dbg & 1024);   // warp-uniform (CTA-uniform); VT_TC_DBG=1024: A/B switch, results valid
        if constexpr (kLoads) {
          if (full) issue(0, std::true_type{});
          else issue(0, std::false_type{});
        }
        mbar_wait(&acc_full[buf], (uint32_t)(seq >> 1) & 1u);
        tc_fence_after();
        if (!(a.dbg & 4)) {
          if (full) {
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) block(cc, std::true_type{});
          } else {
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) block(cc, std::false_type{});
          }
        }
        tc_fence_before();
        mbar_arrive_warp(&acc_empty[buf]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int C, int EM, typename ActT>
int launch_convT_em(const ConvArgs& a, const void* wtc, uint32_t idesc, int grid, cudaStream_t st) {
  constexpr int NA = C == 128 ? 2 : 1;
  constexpr int smem = NA * (C / 64) * kTRA * 128 + kTWst * 128 * 128 + (2 * NA * (C / 64) + 2 * kTWst + 4) * 8 + 16;
  static_assert(smem <= 232448, "shared memory budget exceeded");
  static bool configured = false;
  if (!configured) {
    VT_CUDA_OK(cudaFuncSetAttribute(k_convT_tc<C, EM, ActT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  VT_CUDA_OK(launch_pdl(k_convT_tc<C, EM, ActT>, dim3((unsigned)grid), dim3((kTEpi + 2 + kProdWarps) * 32), (size_t)smem, st, a, reinterpret_cast<const uint8_t*>(wtc), idesc));
  VT_LAUNCHED();
  return VT_OK;
}

template <int C, typename ActT>
int launch_convT_t(const ConvArgs& a, const void* wtc, uint32_t idesc, int grid, cudaStream_t st) {
  int nsnake = 0;
  while (nsnake < 3 && a.act[nsnake].dst && a.act[nsnake].kind == ACT_SNAKE) ++nsnake;
  const bool oact = nsnake == 0 && a.act[0].dst && a.act[0].kind == ACT_LRELU && a.act_from_out;
  for (int s = nsnake + (oact ? 1 : 0); s < 3; ++s) VT_REQUIRE(!a.act[s].dst, "convT_tc: unsupported activation-copy combination");
  const bool r1 = a.res1 != nullptr, r2 = a.res2 != nullptr, out = a.out != nullptr;
  if (!r1 && !r2 && !out && nsnake == 1) return launch_convT_em<C, EM_ACT1, ActT>(a, wtc, idesc, grid, st);
  if (r1 && !r2 && out && !a.out_accum && nsnake == 1) return launch_convT_em<C, EM_RES1 | EM_OUT | EM_ACT1, ActT>(a, wtc, idesc, grid, st);
  if (r1 && r2 && out && !a.out_accum && nsnake == 3)
    return launch_convT_em<C, EM_RES1 | EM_RES2 | EM_OUT | EM_ACT3, ActT>(a, wtc, idesc, grid, st);
  if (r1 && !r2 && out && nsnake == 0 && !oact) return launch_convT_em<C, EM_RES1 | EM_OUT | EM_ACCUM, ActT>(a, wtc, idesc, grid, st);
  if (r1 && !r2 && out && nsnake == 0 && oact)
    return launch_convT_em<C, EM_RES1 | EM_OUT | EM_ACCUM | EM_OACT, ActT>(a, wtc, idesc, grid, st);
  VT_REQUIRE(false, "convT_tc: no compiled epilogue for res1=%d res2=%d out=%d accum=%d snake=%d", (int)r1, (int)r2, (int)out,
             a.out_accum, nsnake);
  return VT_OK;
}

}  // namespace tc

bool convT_tc_supported(const ConvLayer& L) {
  static const bool on = !(getenv("VT_CONVT") && getenv("VT_CONVT")[0] == '0');
  return on && L.w_tc && L.cin == L.cout && (L.cin == 256 || L.cin == 128) && L.stride == 1 && L.out_mul == 1 && (L.k - 1) * L.dil <= 50 && L.pad <= kGap;
}

// `tiles`: tiles of 256 output steps.  The weight image is the one pack_conv_tc builds for the 256-column instance.
int launch_convT_tc(const ConvArgs& a_in, const ConvLayer& L, int act_elem, const void* tiles, int n_tiles, cudaStream_t st) {
  VT_REQUIRE(convT_tc_supported(L) && a_in.in_act && (act_elem == ELEM_F16 || act_elem == ELEM_BF16), "convT_tc: layer %s unsupported",
             L.name.c_str());
  if (n_tiles == 0) return VT_OK;
  ConvArgs a = a_in;
  a.tiles = reinterpret_cast<const ConvTile*>(tiles);
  a.n_tiles = n_tiles;
  static const int dbg = getenv("VT_TC_DBG") ? atoi(getenv("VT_TC_DBG")) : 0;
  a.dbg = dbg;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    VT_CUDA_OK(cudaGetDevice(&dev));
    VT_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  const uint32_t fmt = act_elem == ELEM_F16 ? 0u : 1u;
  // M = 128 channels, N = 256 time steps
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
  if (L.cin == 256)
    return act_elem == ELEM_F16 ? tc::launch_convT_t<256, __half>(a, L.w_tc, idesc, grid, st)
                                : tc::launch_convT_t<256, __nv_bfloat16>(a, L.w_tc, idesc, grid, st);
  return act_elem == ELEM_F16 ? tc::launch_convT_t<128, __half>(a, L.w_tc, idesc, grid, st)
                              : tc::launch_convT_t<128, __nv_bfloat16>(a, L.w_tc, idesc, grid, st);
}

}  // namespace vt
