// Fused ResBlock pair on tcgen05:  x_out = conv2(Snake2(conv1(Snake1(x)))) + x  in ONE kernel
// (upstream hifigan.py ResBlock.forward, one iteration of its dilation loop).
//
// Why: run as two launches the pair moves 1 KB of HBM per time step at C = 64 (operand copy in and
// out of conv1, operand copy + fp32 residual in and fp32 stream + operand copy out of conv2) for
// 2*C*C*k*2 FLOP - below the B200 ridge (210 FLOP/B) for every kernel size at C = 64.  Fused, the only
// HBM traffic of a pair is the fp32 residual stream read once and written once (512 B per step at C = 64).
//
// Per CTA tile (256 conv1 rows = 256 - (k-1) output steps):
//   producers (4 warps)  x fp32 [256 + (k-1)*dil rows] --Snake1--> fp16 A1 tile in shared memory (SWIZZLE_128B)
//   MMA warp   conv1     A1 (taps = row-shifted descriptors) x W1 (bulk-TMA ring)      -> D1 in TMEM
//   mid warps            D1 + b1 --Snake2--> fp16 A2 tile in shared memory; rows outside the sequence -> 0
//   MMA warp   conv2     A2 x W2                                                        -> D2 in TMEM
//   fin warps  (4)       D2 + b2 + x (+ second residual) -> fp32 stream / 1/3-mean accumulate / leaky-ReLU copy
// With NBUF = 2 (C = 64) every buffer (A1, A2, D1, D2) is double buffered and the MMA warp issues
// conv1(i+1) before conv2(i), so the tensor pipe never waits for the mid epilogue; with NBUF = 1
// (C = 128, TMEM and shared memory are full) conv2(i) waits for mid(i) while producers and the fin
// epilogue still overlap.
#include "vt_tc.cuh"

#include <cstdlib>
#include <type_traits>

namespace vt {


namespace tc {

// Residual handling of the fin epilogue.  1: the fin warps preload D2 with x + b2 (+ extras) for the tile that
// uses the buffer next (tcgen05.st; conv2 accumulates on top): no load latency in the epilogue, but two more
// shared-memory transposition passes per block.  0: the residual rows are prefetched into registers two blocks
// ahead and added in the coalesced phase of the output pass: half the shared-memory traffic of the epilogue -
// the L1/shared data pipe is the saturated unit of this kernel.
// Measured (B200, 64 x 500 frames): preload wins wherever the tile is long (k >= 7, C = 128: the weight stream
// makes epilogue load latency very long), register prefetch wins for the short k = 3 tiles at C = 64.

constexpr int kPairRA1 = 312;   // A1 rows: 256 + (k-1)*dil <= 306, multiple of 8
constexpr int kPairRA2 = 272;   // A2 rows: 256 + (k-1) <= 266, multiple of 8

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Warp roles: [0, NEPI) fin epilogue; then (NBUF == 2) NEPI mid-epilogue warps, or (NBUF == 1, "combined") the
// fin warps also run the mid epilogue - conv2(i) waits for mid(i) anyway, and the fin staging can then live in
// the A2 tile, which is idle between conv2(i) and mid(i+1); then the MMA warp, the weight producer, NPROD
// activation-producer warps.
// TR ("transposed"): the MMA computes D^T = W^T X^T - weights are the M = 128 operand, the 256 time steps of the
// tile the N operand (one MMA per K step instead of one per 128-row block: the weight slab is fetched from shared
// memory once per 256 rows, -25 % operand traffic at C = 128).  Accumulators come out as lane = channel, column =
// time step, so a warp's lanes are 32 consecutive channels of ONE time step: the fin epilogue's global accesses
// are coalesced 128-byte lines with no shared-memory transposition (no staging at all), the Snake parameters of
// the mid epilogue are per-thread registers, and mid and fin can be separate warps again.
template <int C, int NBUF, int NEPI, int NPROD, bool TR = false>
struct PairCfg {
  static constexpr bool kCombined = NBUF == 1 && !TR;
  static constexpr int W_MID = kCombined ? 0 : NEPI;
  static constexpr int W_MMA = kCombined ? NEPI : 2 * NEPI;
  // C = 128 has a spare warp slot (22 warps are allocated as 24): the x ring's loader thread gets its own warp there;
  // at C = 64 (20 warps exactly) it shares the weight loader's warp.
  static constexpr bool kSplitLoader = C == 128;
  static constexpr int WARPS = W_MMA + 2 + NPROD + (kSplitLoader ? 1 : 0);   // + MMA warp, loader warp(s)
};

template <typename T> __device__ __forceinline__ unsigned short to_op_bits(float v);
template <> __device__ __forceinline__ unsigned short to_op_bits<__half>(float v) { return __half_as_ushort(__float2half_rn(v)); }
template <> __device__ __forceinline__ unsigned short to_op_bits<__nv_bfloat16>(float v) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}

// WS: the packed weights of this pair carry per-output-channel power-of-two scales (ConvArgs::wscale / PairArgs::wscale1)
// that the epilogues undo.  A layer only gets scaled rows when it needs them (vt_hift.cu scale_weight_rows), so the
// common instance (WS = false) pays nothing: the activation-major epilogues fetch per-channel constants from shared memory
// for every element group, and one more vector there costs 6-9 % of a C = 64 launch (and 40-70 % on the accumulate
// variants, whose fin passes run at the register cap).
//
// NS = 3 ("mean-fused"): the last pairs of the three ResBlocks of a stage as three sub-iterations per tile.  Every role
// but the fin epilogue simply runs NS iterations per tile with the sub's parameters (input stream, weights, Snake
// parameters, kernel size, dilation); conv2 of all subs accumulates in the tile's D2 buffer, the fin epilogue runs once
// per tile with the sum of the residual inputs and biases.  The partial mean (one fp32 stream written and one read per
// ResBlock, loaded synchronously by fin passes on the critical path: +0.15 .. 0.55 ms per launch) never exists.
template <int C, int NBUF, int NA1, int NA2, int W_ST, int NEPI, int NPROD, int NSLAB, int EM, bool kPreload, bool TR, bool WS, typename ActT, int NS = 1>
__global__ void __launch_bounds__(PairCfg<C, NBUF, NEPI, NPROD, TR>::WARPS * 32, 1)
k_pair_tc(const ConvArgs a, const PairArgs p, const uint32_t idesc) {
  using PC = PairCfg<C, NBUF, NEPI, NPROD, TR>;
  static_assert(NS == 1 || (NS == 3 && !WS && !PC::kCombined && (TR || kPreload) && !(EM & (EM_RES2 | EM_ACCUM))),
                "mean-fused launches: unscaled weights, separate mid / fin warps, no other extra stream");
  constexpr int PRM_SUB = 5 * C;                         // al1 | ia1 | b1 | al2 | ia2 of one sub
  constexpr int PRM_FLOATS = NS == 1 ? 8 * C : NS * PRM_SUB;
  auto sub_k = [&](int r) { return (NS == 1 || r == 0) ? p.k : p.more[r - 1].k; };
  auto sub_dil = [&](int r) { return (NS == 1 || r == 0) ? p.dil : p.more[r - 1].dil; };
  auto sub_x = [&](int r) { return (NS == 1 || r == 0) ? p.x_in : p.more[r - 1].x_in; };
  auto sub_h2 = [&](int r) { return (sub_k(r) - 1) / 2; };
  auto sub_h1 = [&](int r) { return (sub_k(r) - 1) * sub_dil(r) / 2; };
  static_assert(!TR || (C == 128 && NBUF == 1 && !kPreload && NEPI == 8), "the transposed variant is built for C = 128");
  constexpr bool kCombined = PC::kCombined;
  constexpr int NMID = NEPI, kFin = NEPI, kProdT = NPROD * 32;
  static_assert(kSwz, "the fused pair kernel assumes the SWIZZLE_128B operand layout");
  constexpr int CB = C / 64;
  constexpr int A1_BYTES = CB * kPairRA1 * 128, A2_BYTES = CB * kPairRA2 * 128, W_BYTES = C * 128;
  constexpr int ACC_COLS = 2 * C;                        // two 128-row M blocks
  constexpr int SKEW = NBUF - 1;
  constexpr int W_MID = PC::W_MID, W_MMA = PC::W_MMA, W_WP = W_MMA + 1, W_AP = W_MMA + 2;
  // fp32 stream slabs of the x ring: kSlabBytes of whole rows each (rows are contiguous in HBM -> one 1-D bulk copy)
  constexpr int kSlabBytes = 8192, SLAB_ROWS = kSlabBytes / (C * 4);
  static_assert(4 * C * NBUF <= 512 && NEPI % 4 == 0 && kProdT % (C / 8) == 0 && SLAB_ROWS % (kProdT / (C / 8)) == 0, "bad configuration");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA1 = smem;
  uint8_t* sA2 = sA1 + NA1 * A1_BYTES;
  uint8_t* sW = sA2 + NA2 * A2_BYTES;
  uint8_t* sX = sW + W_ST * W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + NSLAB * kSlabBytes);
  uint64_t* a1_full = bars;
  uint64_t* a1_empty = a1_full + NA1;
  uint64_t* x_full = a1_empty + NA1;
  uint64_t* x_empty = x_full + NSLAB;
  uint64_t* a2_full = x_empty + NSLAB;
  uint64_t* a2_empty = a2_full + NA2;
  uint64_t* d1_full = a2_empty + NA2;
  uint64_t* d1_empty = d1_full + NBUF;
  uint64_t* d2_full = d1_empty + NBUF;
  uint64_t* d2i_full = d2_full + NBUF;   // D2 free AND preloaded with the residual terms
  uint64_t* w_full = d2i_full + NBUF;
  uint64_t* w_empty = w_full + W_ST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + W_ST);
  float* prm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tmem_slot) + 16);   // al1 | ia1 | b1 | al2 | ia2 | ws1 | ws2 | 1 / ws2
  // fin staging (32 x kStageLd floats per warp): its own region, or the idle A2 tile in combined mode
  float* stage_all = kCombined ? reinterpret_cast<float*>(sA2) : prm + PRM_FLOATS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NA1; ++i) { mbar_init(&a1_full[i], NPROD); mbar_init(&a1_empty[i], 1); }
    for (int i = 0; i < NA2; ++i) { mbar_init(&a2_full[i], NMID); mbar_init(&a2_empty[i], 1); }
    for (int i = 0; i < NSLAB; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], NPROD); }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], NMID);
      mbar_init(&d2_full[i], 1); mbar_init(&d2i_full[i], kFin);
    }
    // weight-ring slots are released by the MMA warps of every CTA that received the multicast copy
    for (int i = 0; i < W_ST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], a.mc ? 2 : 1); }
    fence_barrier_init();
  }
  for (int cr = threadIdx.x; cr < C * NS; cr += blockDim.x) {
    const int r = cr / C, c = cr - r * C;
    const float a1 = (r == 0 ? p.alpha1 : p.more[r - 1].alpha1)[c], a2 = (r == 0 ? p.alpha2 : p.more[r - 1].alpha2)[c];
    float* q = prm + r * PRM_SUB;
    q[c] = a1; q[C + c] = __fdividef(1.0f, a1 + 1e-9f); q[2 * C + c] = (r == 0 ? p.bias1 : p.more[r - 1].bias1)[c];
    q[3 * C + c] = a2; q[4 * C + c] = __fdividef(1.0f, a2 + 1e-9f);
    if constexpr (WS) {
      prm[5 * C + c] = p.wscale1[c];
      const float w2 = a.wscale[c];
      prm[6 * C + c] = w2; prm[7 * C + c] = __frcp_rn(w2);     // shared memory: the fin passes must not wait on global loads
    }
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
  if (!(warp == W_WP && lane == 0)) pdl_wait();   // everything but the (static) weight stream waits for the previous kernel
                                                  // (at C = 64 the x loader is lane 1 of the weight loader's warp: it waits)
  // Weight multicast (a.mc): the two CTAs of a cluster walk the same weight sequence in lockstep, each fetching half
  // of every ring slot for both.  They must run the same number of tiles: an odd tile count is padded with a dummy
  // (the last tile again with n = 0: every store of the epilogues is masked by n).
  if (a.mc) cluster_sync_all();              // the peer's barriers are initialised before anything remote lands on them

  const int n_tiles = a.mc ? (a.n_tiles + 1) & ~1 : a.n_tiles;
  const int n_my = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto get_tile = [&](int i) {
    const int t = (int)blockIdx.x + i * (int)gridDim.x;
    ConvTile tl = a.tiles[t < a.n_tiles ? t : a.n_tiles - 1];
    if (t >= a.n_tiles) tl.n = 0;
    return tl;
  };
  const int n_it = n_my * NS;                  // iterations of every role but fin: (tile, sub) pairs, sub fastest

  constexpr bool kSplitLoader = PC::kSplitLoader;
  constexpr int W_XL = kSplitLoader ? W_AP + NPROD : W_WP;    // x loader: its own warp, or lane 1 of the weight loader's
  constexpr int kXlLane = kSplitLoader ? 0 : 1;
  if (warp >= W_AP && warp < W_AP + NPROD) {
    // ---------------- producers: fp32 slabs of the x ring -> Snake1 -> fp16 A1 tile (SWIZZLE_128B).  All global
    // latency is taken by the bulk-copy engine; this loop is shared-memory to shared-memory.  A thread owns the
    // channels [4*ch, 4*ch+4) and [C/2 + 4*ch, +4) (two conflict-free 16-byte reads per row) for every row it
    // touches, so its Snake parameters stay in registers.
    constexpr int QPR = C / 8;                     // threads per row
    constexpr int RPP = kProdT / QPR;              // rows per pass of all producer threads
    const int pt = threadIdx.x - W_AP * 32;
    const int ch = pt % QPR, r_raw = pt / QPR;
    // C = 64: swap bits 0 and 2 of the row index.  An 8-byte store is served per HALF warp (16 lanes x 8 B = 128 B = all
    // 32 banks once if the lanes hit 8 distinct 16-byte chunks); a half warp holds two rows, each writing chunks
    // (0..3) ^ (row & 7) - two rows on the same side of (row & 4) collide (ncu source view: every A1 store took twice
    // its ideal wavefronts, 6-10 % of all shared-memory wavefronts of the C = 64 launches), rows 4 apart do not.
    // C = 128: a half warp is one whole 128-byte row either way (bits 1 and 2 swapped as before).
#ifdef VT_OLD_ROWPERM          // A/B build switch (tools/ab_build.sh oldperm -DVT_OLD_ROWPERM)
    const int r_in = (r_raw & ~6) | ((r_raw & 2) << 1) | ((r_raw & 4) >> 1);
#else
    const int r_in = C == 64 ? ((r_raw & ~5) | ((r_raw & 1) << 2) | ((r_raw & 4) >> 2))
                             : ((r_raw & ~6) | ((r_raw & 2) << 1) | ((r_raw & 4) >> 1));
#endif
    const int cA = 4 * ch, cB = C / 2 + 4 * ch;    // first channel of the two pieces
    float alA[4], iaA[4], alB[4], iaB[4];
    auto load_snake1 = [&](int r) {
      const float* q = prm + r * PRM_SUB;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        alA[e] = q[cA + e]; iaA[e] = q[C + cA + e];
        alB[e] = q[cB + e]; iaB[e] = q[C + cB + e];
      }
    };
    load_snake1(0);
    // byte offsets of the two pieces inside an A1 row block: 64-channel block, 16-byte chunk, 8-byte half
    const uint32_t blkA = (uint32_t)(cA >> 6) * (uint32_t)(kPairRA1 * 128), blkB = (uint32_t)(cB >> 6) * (uint32_t)(kPairRA1 * 128);
    const uint32_t chkA = (uint32_t)((cA & 63) >> 3), chkB = (uint32_t)((cB & 63) >> 3);
    const uint32_t halfA = (uint32_t)((cA & 7) >> 2) * 8u, halfB = (uint32_t)((cB & 7) >> 2) * 8u;
    const uint32_t a1_base = smem_u32(sA1);
    uint32_t xs = 0, xph = 0;
    for (int i = 0; i < n_it; ++i) {
      const int b1 = i % NA1;
      if constexpr (NS > 1) load_snake1(i % NS);
      const int R1 = 256 + 2 * sub_h1(i % NS);
      const int n_slab = (R1 + SLAB_ROWS - 1) / SLAB_ROWS;
      mbar_wait(&a1_empty[b1], ((uint32_t)(i / NA1) & 1u) ^ 1u);
      if (pt == 0) trace_ev(a.trace, i, 0);
      const uint32_t a1 = a1_base + (uint32_t)(b1 * A1_BYTES);
      long long x_wait = 0;
      for (int sl = 0; sl < n_slab; ++sl) {
        if (a.trace) {
          const long long tw = clock64();
          mbar_wait(&x_full[xs], xph);
          x_wait += clock64() - tw;
        } else {
          mbar_wait(&x_full[xs], xph);
        }
        const uint32_t xsrc = smem_u32(sX + xs * kSlabBytes);
        // all shared-memory reads of this slab first (the volatile asm statements keep program order, so
        // interleaving loads and stores task by task would serialise the tasks), then convert and store
        constexpr int TPS = SLAB_ROWS / RPP;          // tasks per thread per slab
        float4 va[TPS], vb[TPS];
#pragma unroll
        for (int t = 0; t < TPS; ++t) {
          const int rr = r_in + t * RPP;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(va[t].x), "=f"(va[t].y), "=f"(va[t].z), "=f"(va[t].w)
                       : "r"(xsrc + (uint32_t)(rr * C * 4 + cA * 4)));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vb[t].x), "=f"(vb[t].y), "=f"(vb[t].z), "=f"(vb[t].w)
                       : "r"(xsrc + (uint32_t)(rr * C * 4 + cB * 4)));
        }
#pragma unroll
        for (int t = 0; t < TPS; ++t) {
          const int r = sl * SLAB_ROWS + r_in + t * RPP;
          if (r < R1) {
            float ya[4] = {snake_f(va[t].x, alA[0], iaA[0]), snake_f(va[t].y, alA[1], iaA[1]), snake_f(va[t].z, alA[2], iaA[2]),
                           snake_f(va[t].w, alA[3], iaA[3])};
            float yb[4] = {snake_f(vb[t].x, alB[0], iaB[0]), snake_f(vb[t].y, alB[1], iaB[1]), snake_f(vb[t].z, alB[2], iaB[2]),
                           snake_f(vb[t].w, alB[3], iaB[3])};
            const uint2 pa = Pack4<ActT>::pack(ya), pb = Pack4<ActT>::pack(yb);
            const uint32_t rowb = a1 + (uint32_t)r * 128u;
            const uint32_t swz = (uint32_t)(r & 7);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowb + blkA + ((chkA ^ swz) << 4) + halfA), "r"(pa.x), "r"(pa.y) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowb + blkB + ((chkB ^ swz) << 4) + halfB), "r"(pb.x), "r"(pb.y) : "memory");
          }
        }
        mbar_arrive_warp(&x_empty[xs]);                  // the slab has been read: the loader may refill it
        if (++xs == (uint32_t)NSLAB) { xs = 0; xph ^= 1u; }
      }
      fence_proxy_async();
      if (pt == 0) trace_ev(a.trace, i, 1);
      if (pt == 0 && a.trace && blockIdx.x == 0 && i < kTraceTiles) a.trace[i * kTraceEvents + 12] = x_wait;
      mbar_arrive_warp(&a1_full[b1]);
    }
  } else if (warp == W_XL || warp == W_WP) {
    // ---------------- loaders: one THREAD per ring, each in its own blocking loop (a thread that polled both rings
    // delayed weight copies behind tile-table loads and x slabs).  C = 128 has a spare warp slot for the x loader; at
    // C = 64 (20 warps exactly) it is lane 1 of the weight loader's warp - diverged lanes make progress independently.
    if (warp == W_XL && lane == kXlLane) {
      // x ring: the tiles' fp32 rows (with halo), slab by slab (whole rows are contiguous in HBM -> 1-D bulk copies)
      uint32_t xs = 0, xph = 0;
      ConvTile tl = n_my > 0 ? get_tile(0) : ConvTile{};
      for (int xi = 0; xi < n_it; ++xi) {
        const int r = xi % NS, H1 = sub_h1(r), H2 = sub_h2(r);
        const int R1 = 256 + 2 * H1;
        const float* xsrc = sub_x(r) + (tl.in_row0 + tl.q0 - H1 - H2) * (long long)C;
        if (r == NS - 1 && xi + 1 < n_it) tl = get_tile(xi / NS + 1);   // next tile's entry: off the critical path
        for (int xr = 0; xr < R1; xr += SLAB_ROWS) {
          const int rows = R1 - xr < SLAB_ROWS ? R1 - xr : SLAB_ROWS;
          mbar_wait(&x_empty[xs], xph ^ 1u);
          if (a.dbg & 2) mbar_arrive(&x_full[xs]);
          else {
            mbar_arrive_expect_tx(&x_full[xs], (uint32_t)(rows * C * 4));
            bulk_g2s(sX + xs * kSlabBytes, xsrc + (long long)xr * C, (uint32_t)(rows * C * 4), &x_full[xs]);
          }
          if (++xs == (uint32_t)NSLAB) { xs = 0; xph ^= 1u; }
        }
      }
    } else if (warp == W_WP && lane == 0) {
      // weight ring: chunk order mirrors the MMA issue order
      const uint32_t mc_rank = a.mc ? cluster_ctarank() : 0u;
      uint32_t ws = 0, wph = 0;
      for (int s = 0; s < n_it + SKEW; ++s)
        for (int pass = 0; pass < 2; ++pass) {
          if (pass == 0 ? s >= n_it : s < SKEW) continue;
          const int r = (pass == 0 ? s : s - SKEW) % NS;
          const uint8_t* wsrc = (NS == 1 || r == 0) ? (pass == 0 ? p.w1 : p.w2) : (pass == 0 ? p.more[r - 1].w1 : p.more[r - 1].w2);
          const int nchunks = sub_k(r) * CB;
          for (int w_c = 0; w_c < nchunks; ++w_c) {
            mbar_wait(&w_empty[ws], wph ^ 1u);
            if (a.dbg & 1) mbar_arrive(&w_full[ws]);
            else {
              mbar_arrive_expect_tx(&w_full[ws], W_BYTES);
              if (a.mc) {
                const uint32_t hoff = mc_rank * (uint32_t)(W_BYTES / 2);
                bulk_g2s_mc(sW + ws * W_BYTES + hoff, wsrc + (size_t)w_c * W_BYTES + hoff, W_BYTES / 2, &w_full[ws], (uint16_t)3);
              } else {
                bulk_g2s(sW + ws * W_BYTES, wsrc + (size_t)w_c * W_BYTES, W_BYTES, &w_full[ws]);
              }
            }
            if (++ws == (uint32_t)W_ST) { ws = 0; wph ^= 1u; }
          }
        }
      if (a.mc) {
        // the peer's last slot releases arrive on THIS CTA's barriers: take them before the CTA may exit
        for (int q = 0; q < W_ST; ++q) {
          mbar_wait(&w_empty[ws], wph ^ 1u);
          if (++ws == (uint32_t)W_ST) { ws = 0; wph ^= 1u; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ---------------- MMA issuer: conv1(s) then conv2(s - SKEW)
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a1_lo0 = ((smem_u32(sA1) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t a2_lo0 = ((smem_u32(sA2) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t w_lo0 = ((smem_u32(sW) >> 4) & 0x3FFFu) | (1u << 16);
    const bool mma_on = !(a.dbg & 16);
    uint32_t ws = 0, wph = 0;
    for (int s = 0; s < n_it + SKEW; ++s)
      for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0 ? s >= n_it : s < SKEW) continue;
        const int i = pass == 0 ? s : s - SKEW;
        const int r = i % NS;                                       // sub-pair of the tile
        // D1 is double buffered by iteration; D2 belongs to the TILE: conv2 of every sub accumulates in it
        const int ib = pass == 0 ? i : i / NS;
        const int b = ib % NBUF;
        const uint32_t u = (uint32_t)(ib / NBUF);
        const int bs = pass == 0 ? i % NA1 : i % NA2;                         // shared-memory operand buffer
        const uint32_t us = (uint32_t)(pass == 0 ? i / NA1 : i / NA2);
        uint64_t* src_full = pass == 0 ? &a1_full[bs] : &a2_full[bs];
        uint64_t* src_empty = pass == 0 ? &a1_empty[bs] : &a2_empty[bs];
        uint64_t* dst_full = pass == 0 ? &d1_full[b] : &d2_full[b];
        // conv1 needs D1 drained by the mid epilogue; conv2 needs D2 preloaded by the fin warps (x + b2 + ...)
        if (pass == 0) mbar_wait(&d1_empty[b], (u & 1u) ^ 1u);
        else if (r == 0) {
          if (kPreload) mbar_wait(&d2i_full[b], u & 1u);
          else mbar_wait(&d2i_full[b], (u & 1u) ^ 1u);     // plain "D2 drained" barrier in register-prefetch mode
        }
        mbar_wait(src_full, us & 1u);
        tc_fence_after();
        if (lane == 0) trace_ev(a.trace, i, 6 + 2 * pass);
        const uint32_t d0 = tmem_base + (uint32_t)((pass * NBUF + b) * ACC_COLS);
        const uint32_t a_tile = pass == 0 ? a1_lo0 + (uint32_t)bs * (uint32_t)(A1_BYTES >> 4) : a2_lo0 + (uint32_t)bs * (uint32_t)(A2_BYTES >> 4);
        const uint32_t blk16 = (uint32_t)(pass == 0 ? kPairRA1 : kPairRA2) * 8u;   // 64-channel block stride, 16-byte units
        const uint32_t tap16 = (uint32_t)(pass == 0 ? sub_dil(r) : 1) * 8u;         // one tap = dil rows of 128 B
        uint32_t acc = (pass == 1 && (kPreload || r > 0)) ? 1u : 0u;
        long long w_wait = 0;
        const int k_r = sub_k(r);
        for (int j = 0; j < k_r; ++j) {
          uint32_t a_chunk = a_tile + (uint32_t)j * tap16;
#pragma unroll 1
          for (int cb = 0; cb < CB; ++cb, a_chunk += blk16) {
            if (a.trace) {
              const long long tw = clock64();
              mbar_wait(&w_full[ws], wph);
              w_wait += clock64() - tw;
            } else {
              mbar_wait(&w_full[ws], wph);
            }
            tc_fence_after();
            if (elect_one()) {
              const uint32_t b_lo = w_lo0 + ws * (uint32_t)(W_BYTES >> 4);
              if (mma_on) {
                if constexpr (TR) {
                  // A operand = weight slab (M = C output channels), B operand = 256 rows of the activation tile
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_f16_lh(d0, b_lo + (uint32_t)(ks * 2), a_chunk + (uint32_t)(ks * 2), kDescHi, idesc, ks == 0 ? acc : 1u);
                } else {
#pragma unroll
                  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                      umma_f16_lh(d0 + (uint32_t)(mb * C), a_chunk + (uint32_t)(mb * 1024 + ks * 2), b_lo + (uint32_t)(ks * 2), kDescHi,
                                  idesc, ks == 0 ? acc : 1u);
                }
              }
              if (a.mc) umma_commit_mc(&w_empty[ws], (uint16_t)3);
              else umma_commit(&w_empty[ws]);
            }
            __syncwarp();
            acc = 1u;
            if (++ws == (uint32_t)W_ST) { ws = 0; wph ^= 1u; }
          }
        }
        if (elect_one()) {
          umma_commit(src_empty);
          if (pass == 0 || r == NS - 1) umma_commit(dst_full);      // D2 is complete after the last sub's conv2
        }
        __syncwarp();
        if (lane == 0) trace_ev(a.trace, i, 7 + 2 * pass);
        if (lane == 0 && a.trace && blockIdx.x == 0 && i < kTraceTiles) a.trace[i * kTraceEvents + 10 + pass] = w_wait;
      }
  } else if constexpr (TR) {
    // ---------------- transposed epilogues: lane = output channel (TMEM lane), TMEM column = time step of the tile.
    // Warps [0, NEPI) fin, [NEPI, 2 NEPI) mid; warp & 3 = channel quarter, the two warps of a quarter split the
    // 256 time steps in halves.
    const bool is_fin = warp < W_MID;
    const int ew = is_fin ? warp : warp - W_MID;
    const int quarter = warp & 3, half = ew >> 2;
    const int c = quarter * 32 + lane;                                   // this thread's channel
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    if (!is_fin) {
      float b1 = prm[2 * C + c], al2 = prm[3 * C + c], ia2 = prm[4 * C + c];
      const float ws1 = WS ? prm[5 * C + c] : 1.0f;
      // A2 element (row r, channel c): 64-channel block, 16-byte chunk XOR-swizzled by the row, 2 bytes inside
      const uint32_t coff = (uint32_t)(c >> 6) * (uint32_t)(kPairRA2 * 128) + (uint32_t)((c & 7) * 2);
      const uint32_t chunk = (uint32_t)((c & 63) >> 3);
      ConvTile tile = n_my > 0 ? get_tile(0) : ConvTile{};
      for (int i = 0; i < n_it; ++i) {
        const int r = i % NS;
        if (NS == 1 || r == 0) tile = get_tile(i / NS);
        if constexpr (NS > 1) {
          const float* q = prm + r * PRM_SUB;
          b1 = q[2 * C + c]; al2 = q[3 * C + c]; ia2 = q[4 * C + c];
        }
        const int H2 = sub_h2(r);
        mbar_wait(&d1_full[0], (uint32_t)i & 1u);
        mbar_wait(&a2_empty[0], ((uint32_t)i & 1u) ^ 1u);
        tc_fence_after();
        if (ew == 0 && lane == 0) trace_ev(a.trace, i, 2);
        const uint32_t dst0 = smem_u32(sA2) + coff;
#pragma unroll 1
        for (int cc = 0; cc < 4 && !(a.dbg & 4); ++cc) {
          const int col0 = half * 128 + cc * 32;
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_sel + (uint32_t)col0, v);
          tmem_ld_wait();
          const int p0 = tile.q0 - H2 + col0;                              // sequence position of the block's first row
          if (p0 >= 0 && p0 + 32 <= tile.out_len && !(a.dbg & 1024)) {
            // the whole block lies inside the sequence (all blocks but those at a sequence's ends): no per-row test,
            // and (col0 & 7) == 0 makes the swizzle phase of row col0 + t a compile-time function of t.  The mid pass is
            // on the conv1 -> mid -> conv2 chain that bounds the k >= 7 tiles.
            const uint32_t base = dst0 + (uint32_t)col0 * 128u;
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const float y = snake_f(fmaf(__uint_as_float(v[t]), ws1, b1), al2, ia2);
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(base + ((chunk ^ (uint32_t)(t & 7)) << 4) + (uint32_t)(t * 128)),
                           "h"(to_op_bits<ActT>(y)) : "memory");
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const int row = col0 + t;                                      // conv1 output row of the tile = A2 row
              const int pseq = tile.q0 - H2 + row;
              const bool valid = pseq >= 0 && pseq < tile.out_len;          // warp-uniform
              const float y = valid ? snake_f(fmaf(__uint_as_float(v[t]), ws1, b1), al2, ia2) : 0.0f;
              const uint32_t addr = dst0 + (uint32_t)row * 128u + ((chunk ^ (uint32_t)(row & 7)) << 4);
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(to_op_bits<ActT>(y)) : "memory");
            }
          }
        }
        tc_fence_before();
        fence_proxy_async();
        if (ew == 0 && lane == 0) trace_ev(a.trace, i, 3);
        mbar_arrive_warp(&d1_empty[0]);
        mbar_arrive_warp(&a2_full[0]);
      }
    } else {
      constexpr bool kRes2 = (EM & EM_RES2) != 0, kAccum = (EM & EM_ACCUM) != 0;
      float b2 = a.bias[c];
      if constexpr (NS > 1) b2 += p.more[0].bias2[c] + p.more[1].bias2[c];
      const float ws2 = WS ? a.wscale[c] : 1.0f;
      const bool accum = kAccum && a.out_accum;
      const float inv = 1.0f / a.out_scale;
      const uint32_t d2 = tmem_base + lane_sel + (uint32_t)ACC_COLS;
      // Residual rows are prefetched into registers one 32-step block ahead (the first block of a tile before the
      // d2_full wait): 32 coalesced 128-byte requests in flight per warp, lane = channel on both sides, so neither
      // the loads nor the stores need a transposition.  Unconditional loads from clamped rows (a branch per load
      // would serialise them).
      float x[32];
      // only (first element of this thread's column, rows) of a tile are kept in registers
      auto tile_base = [&](int i, int& n) {
        const ConvTile tl = get_tile(i);
        n = tl.n;
        return (tl.out_row0 + tl.q0) * (long long)C + c;
      };
      // A block whose 32 time steps are all output steps of the tile (every block but a tile's last one or two) takes
      // the "full" paths below: no index clamps, no per-element predicates, one base pointer plus compile-time offsets.
      auto issue = [&](long long base, int n, int cc) {
        const int col0 = half * 128 + cc * 32, last = n - 1;
        if (col0 + 32 <= n && !(a.dbg & 1024)) {
          const float* xr = a.res1 + base + (long long)col0 * C;
#pragma unroll
          for (int t = 0; t < 32; ++t) x[t] = (a.dbg & 32) ? 0.0f : __ldg(xr + t * C);
        } else {
          const float* xr = a.res1 + base;
#pragma unroll
          for (int t = 0; t < 32; ++t) x[t] = (a.dbg & 32) ? 0.0f : __ldg(xr + (long long)(col0 + t < last ? col0 + t : last) * C);
        }
      };
      int n_cur = 1;
      long long obase = n_my > 0 ? tile_base(0, n_cur) : 0;
      if (n_my > 0) issue(obase, n_cur, 0);
      for (int i = 0; i < n_my; ++i) {
        mbar_wait(&d2_full[0], (uint32_t)i & 1u);
        tc_fence_after();
        if (warp == 0 && lane == 0) trace_ev(a.trace, i * NS + NS - 1, 4);
        struct { int n; } tile{n_cur};
        auto block = [&](int cc, auto full_tag) {
          constexpr bool kFull = decltype(full_tag)::value;
          const int col0 = half * 128 + cc * 32, last = tile.n - 1;
          const long long cbase = obase + (long long)col0 * C;
          auto at = [&](int t) { return kFull ? cbase + t * C : obase + (long long)(col0 + t < last ? col0 + t : last) * C; };
          if constexpr (kRes2) {
#pragma unroll
            for (int t = 0; t < 32; ++t) x[t] += __ldg(a.res2 + at(t));
          }
          if (accum) {
#pragma unroll
            for (int t = 0; t < 32; ++t) x[t] = fmaf(__ldg(a.out + at(t)), inv, x[t]);
          }
          if constexpr (NS > 1) {                          // the other subs' residual inputs (pair inputs share the row map)
#pragma unroll
            for (int q = 0; q < NS - 1; ++q) {
              const float* xq = p.more[q].x_in;
#pragma unroll
              for (int t = 0; t < 32; ++t) x[t] += __ldg(xq + at(t));
            }
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {                 // 16 accumulator columns at a time: x[32] stays live
            uint32_t v[16];
            tmem_ld16(d2 + (uint32_t)(col0 + hh * 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const int row = col0 + hh * 16 + t;
              if (kFull || row < tile.n) {
                const float o = fmaf(__uint_as_float(v[t]), ws2, b2 + x[hh * 16 + t]) * a.out_scale;
                const long long idx = kFull ? cbase + (hh * 16 + t) * C : obase + (long long)row * C;
                if (!(a.dbg & 64)) a.out[idx] = o;
                if constexpr ((EM & EM_OACT) != 0) {
                  const float sl = a.act[0].slope;
                  reinterpret_cast<unsigned short*>(a.act[0].dst)[idx] = to_op_bits<ActT>(o > 0.f ? o : o * sl);
                }
              }
            }
          }
        };
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          if (!(a.dbg & 4)) {
            if (half * 128 + cc * 32 + 32 <= tile.n && !(a.dbg & 1024)) block(cc, std::true_type{});
            else block(cc, std::false_type{});
          }
          // next block of this tile, or the first block of the next one (in flight during the d2_full wait)
          if (cc < 3) issue(obase, n_cur, cc + 1);
          else if (i + 1 < n_my) {
            obase = tile_base(i + 1, n_cur);
            issue(obase, n_cur, 0);
          }
        }
        tc_fence_before();
        if (warp == 0 && lane == 0) trace_ev(a.trace, i * NS + NS - 1, 5);
        mbar_arrive_warp(&d2i_full[0]);
      }
    }
  } else {
    // ---------------- epilogue warps.  mid: D1 + b1 -> Snake2 -> fp16 A2 rows (zero outside the sequence);
    // fin: D2 -> fp32 stream, then the residual terms of the tile that uses D2 next are preloaded into it.
    const bool do_mid = kCombined || warp >= W_MID;
    const bool do_fin = kCombined || warp < W_MID;
    const int ew = kCombined ? warp : (warp >= W_MID ? warp - W_MID : warp);   // index inside the role group
    const int quarter = warp & 3, grp = ew >> 2;
    constexpr int NBLK = 2 * (C / 32);
    float* stage = stage_all + ew * (32 * kStageLd);
    // ---- fin helpers.  A block = 32 rows (this warp's TMEM lanes of one M block) x 32 columns.  Global memory
    // is touched as 4 rows x 128 contiguous bytes per instruction (lane -> (rsub, sub)); TMEM wants lane = row;
    // the shared-memory stage transposes between the two in both directions.
    const int sub = lane & 7, rsub = lane >> 3;
    constexpr bool kRes2 = (EM & EM_RES2) != 0, kAccum = (EM & EM_ACCUM) != 0;
    auto blk_geom = [&](const ConvTile& tl, int blk, int& mb, int& c0, long long& idx0, int& nvalid) {
      mb = blk / (C / 32);
      c0 = (blk - mb * (C / 32)) * 32;
      const int row0 = mb * 128 + quarter * 32;
      idx0 = (tl.out_row0 + tl.q0 + row0 + rsub) * (long long)C + c0 + sub * 4;
      nvalid = tl.n - row0 - rsub;                 // rows 4i + rsub of the block are valid while 4i < nvalid
    };
    // residual rows of a FUTURE tile -> registers (asynchronous: nothing below depends on them until fin_store)
    auto fin_issue = [&](const ConvTile& tl, int blk, float4 (&pre)[8]) {
      int mb, c0, nvalid; long long idx0;
      blk_geom(tl, blk, mb, c0, idx0, nvalid);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        pre[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * q < nvalid && !(a.dbg & 32)) pre[q] = ldg_f4(a.res1 + idx0 + (long long)q * 4 * C);
      }
    };
    // x + b2 (+ second residual) (+ previous partial mean / scale) -> D2[b]: conv2 accumulates on top of it
    auto fin_store = [&](const ConvTile& tl, int blk, int b, float4 (&pre)[8]) {
      int mb, c0, nvalid; long long idx0;
      blk_geom(tl, blk, mb, c0, idx0, nvalid);
      float4 bias = *reinterpret_cast<const float4*>(a.bias + c0 + sub * 4);
      if constexpr (NS > 1) {
#pragma unroll
        for (int q = 0; q < NS - 1; ++q) {
          const float4 bq = *reinterpret_cast<const float4*>(p.more[q].bias2 + c0 + sub * 4);
          bias.x += bq.x; bias.y += bq.y; bias.z += bq.z; bias.w += bq.w;
        }
      }
      // conv2 accumulates s_c * (W2 a) on top of the preload, so the preload is s_c * (x + b2 + ...): exact (power of two)
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
      if constexpr (WS) sc = *reinterpret_cast<const float4*>(prm + 7 * C + c0 + sub * 4);
      const bool accum = kAccum && a.out_accum;
      const float inv = 1.0f / a.out_scale;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * q < nvalid) {
          t = make_float4(pre[q].x + bias.x, pre[q].y + bias.y, pre[q].z + bias.z, pre[q].w + bias.w);
          const long long idx = idx0 + (long long)q * 4 * C;
          if constexpr (kRes2) {
            const float4 r2 = ldg_f4(a.res2 + idx);
            t.x += r2.x; t.y += r2.y; t.z += r2.z; t.w += r2.w;
          }
          if (accum) {
            const float4 pv = *reinterpret_cast<const float4*>(a.out + idx);
            t.x = fmaf(pv.x, inv, t.x); t.y = fmaf(pv.y, inv, t.y); t.z = fmaf(pv.z, inv, t.z); t.w = fmaf(pv.w, inv, t.w);
          }
          if constexpr (NS > 1) {                            // the other subs' residual inputs
#pragma unroll
            for (int m = 0; m < NS - 1; ++m) {
              const float4 xm = ldg_f4(p.more[m].x_in + idx);
              t.x += xm.x; t.y += xm.y; t.z += xm.z; t.w += xm.w;
            }
          }
          if constexpr (WS) { t.x *= sc.x; t.y *= sc.y; t.z *= sc.z; t.w *= sc.w; }
        }
        *reinterpret_cast<float4*>(stage + (q * 4 + rsub) * kStageLd + sub * 4) = t;
      }
      __syncwarp();
      uint32_t v[32];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(stage + lane * kStageLd + g * 4);
        v[4 * g] = __float_as_uint(t.x); v[4 * g + 1] = __float_as_uint(t.y);
        v[4 * g + 2] = __float_as_uint(t.z); v[4 * g + 3] = __float_as_uint(t.w);
      }
      __syncwarp();
      tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((NBUF + b) * ACC_COLS + mb * C + c0), v);
    };
    // D2[b] -> out = acc * scale (+ leaky-ReLU operand copy): no global reads on this path
    auto fin_out = [&](const ConvTile& tl, int blk, int b) {
      int mb, c0, nvalid; long long idx0;
      blk_geom(tl, blk, mb, c0, idx0, nvalid);
      // 8-column loads: the prefetched residual registers of fin_issue are live across this function
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        uint32_t v[8];
        tmem_ld8(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((NBUF + b) * ACC_COLS + mb * C + c0 + hh * 8), v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g)
          *reinterpret_cast<float4*>(stage + lane * kStageLd + hh * 8 + g * 4) =
              make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                          __uint_as_float(v[4 * g + 3]));
      }
      __syncwarp();
      if (!(a.dbg & 8)) {
        float4 osc = make_float4(a.out_scale, a.out_scale, a.out_scale, a.out_scale);
        if constexpr (WS) {
          const float4 wsq = *reinterpret_cast<const float4*>(prm + 6 * C + c0 + sub * 4);
          osc = make_float4(wsq.x * a.out_scale, wsq.y * a.out_scale, wsq.z * a.out_scale, wsq.w * a.out_scale);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (4 * q >= nvalid) continue;
          const long long idx = idx0 + (long long)q * 4 * C;
          const float4 acc = *reinterpret_cast<const float4*>(stage + (q * 4 + rsub) * kStageLd + sub * 4);
          const float4 o = make_float4(acc.x * osc.x, acc.y * osc.y, acc.z * osc.z, acc.w * osc.w);
          if (!(a.dbg & 64)) *reinterpret_cast<float4*>(a.out + idx) = o;
          if constexpr ((EM & EM_OACT) != 0) {
            const float sl = a.act[0].slope;
            float y[4] = {o.x > 0.f ? o.x : o.x * sl, o.y > 0.f ? o.y : o.y * sl, o.z > 0.f ? o.z : o.z * sl,
                          o.w > 0.f ? o.w : o.w * sl};
            *reinterpret_cast<uint2*>(reinterpret_cast<ActT*>(a.act[0].dst) + idx) = Pack4<ActT>::pack(y);
          }
        }
      }
      __syncwarp();
    };
    // ---- register-prefetch mode: block sequence n = tile * BPW + k of this warp; residual rows of block n + PF are
    // requested right after block n has been consumed
    constexpr int BPW = NBLK / (NEPI / 4), PF = C == 64 ? 1 : 2;    // 96 registers at C = 64 hold one block in flight
    static_assert(BPW % PF == 0, "prefetch slots must divide the blocks per warp");
    auto pf_issue = [&](int n, float4 (&pr)[8]) {
      const int ti = n / BPW;
      if (ti >= n_my) return;
      const ConvTile tl = get_tile(ti);
      int mb, c0, nvalid; long long idx0;
      blk_geom(tl, grp + (n - ti * BPW) * (NEPI / 4), mb, c0, idx0, nvalid);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        pr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * q < nvalid && !(a.dbg & 32)) pr[q] = ldg_f4(a.res1 + idx0 + (long long)q * 4 * C);
      }
    };
    auto fin_out_res = [&](const ConvTile& tl, int blk, int b, float4 (&pr)[8]) {
      int mb, c0, nvalid; long long idx0;
      blk_geom(tl, blk, mb, c0, idx0, nvalid);
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        uint32_t v[8];
        tmem_ld8(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((NBUF + b) * ACC_COLS + mb * C + c0 + hh * 8), v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g)
          *reinterpret_cast<float4*>(stage + lane * kStageLd + hh * 8 + g * 4) =
              make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                          __uint_as_float(v[4 * g + 3]));
      }
      __syncwarp();
      if (!(a.dbg & 8)) {
        const float4 bias = *reinterpret_cast<const float4*>(a.bias + c0 + sub * 4);
        const float4 wsq = WS ? *reinterpret_cast<const float4*>(prm + 6 * C + c0 + sub * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
        const bool accum = kAccum && a.out_accum;
        const float inv = 1.0f / a.out_scale;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (4 * q >= nvalid) continue;
          const long long idx = idx0 + (long long)q * 4 * C;
          const float4 acc = *reinterpret_cast<const float4*>(stage + (q * 4 + rsub) * kStageLd + sub * 4);
          float4 t = make_float4(fmaf(acc.x, wsq.x, bias.x + pr[q].x), fmaf(acc.y, wsq.y, bias.y + pr[q].y),
                                 fmaf(acc.z, wsq.z, bias.z + pr[q].z), fmaf(acc.w, wsq.w, bias.w + pr[q].w));
          if constexpr (kRes2) {
            const float4 r2 = ldg_f4(a.res2 + idx);
            t.x += r2.x; t.y += r2.y; t.z += r2.z; t.w += r2.w;
          }
          if (accum) {
            const float4 pv = *reinterpret_cast<const float4*>(a.out + idx);
            t.x = fmaf(pv.x, inv, t.x); t.y = fmaf(pv.y, inv, t.y); t.z = fmaf(pv.z, inv, t.z); t.w = fmaf(pv.w, inv, t.w);
          }
          const float4 o = make_float4(t.x * a.out_scale, t.y * a.out_scale, t.z * a.out_scale, t.w * a.out_scale);
          if (!(a.dbg & 64)) *reinterpret_cast<float4*>(a.out + idx) = o;
          if constexpr ((EM & EM_OACT) != 0) {
            const float sl = a.act[0].slope;
            float y[4] = {o.x > 0.f ? o.x : o.x * sl, o.y > 0.f ? o.y : o.y * sl, o.z > 0.f ? o.z : o.z * sl,
                          o.w > 0.f ? o.w : o.w * sl};
            *reinterpret_cast<uint2*>(reinterpret_cast<ActT*>(a.act[0].dst) + idx) = Pack4<ActT>::pack(y);
          }
        }
      }
      __syncwarp();
    };
    float4 pfr[PF][8];
    if (do_fin && !kPreload) {
#pragma unroll
      for (int k = 0; k < PF; ++k) pf_issue(k, pfr[k]);
    }
    // prologue: preload D2 for the first NBUF tiles
    if (do_fin && kPreload) {
      for (int j = 0; j < NBUF && j < n_my; ++j) {
        const ConvTile tl = get_tile(j);
#pragma unroll 1
        for (int blk = grp; blk < NBLK; blk += NEPI / 4) {
          float4 pre[8];
          fin_issue(tl, blk, pre);
          fin_store(tl, blk, j, pre);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_warp(&d2i_full[j]);
      }
    }
    // one loop over (tile, sub) iterations: the mid part runs every iteration, the fin part once per tile (after the
    // last sub); NS > 1 has separate mid and fin warps, so each warp only sees its own part
    for (int it = 0; it < n_it; ++it) {
      const int sr = it % NS;                              // sub-pair
      const ConvTile tile = get_tile(it / NS);
      if (do_mid) {
        const int i = it, b = i % NBUF, H2 = sub_h2(sr);
        const uint32_t u = (uint32_t)(i / NBUF);
        const float* prs = prm + sr * PRM_SUB;
        const int b2 = i % NA2;
        mbar_wait(&d1_full[b], u & 1u);
        mbar_wait(&a2_empty[b2], ((uint32_t)(i / NA2) & 1u) ^ 1u);
        tc_fence_after();
        if (ew == 0 && lane == 0 && (kCombined || warp == W_MID)) trace_ev(a.trace, i, 2);
        const uint32_t dst0 = smem_u32(sA2 + b2 * A2_BYTES);
#pragma unroll 1
        for (int blk = grp; blk < NBLK && !(a.dbg & 4); blk += NMID / 4) {
          const int mb = blk / (C / 32), c0 = (blk - mb * (C / 32)) * 32;
          const int row = mb * 128 + quarter * 32 + lane;              // A2 row = conv1 output row of the tile
          const int pseq = tile.q0 - H2 + row;                         // step inside the sequence
          const bool valid = pseq >= 0 && pseq < tile.out_len;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * ACC_COLS + mb * C + c0), v);
          tmem_ld_wait();
          const uint32_t rowaddr = dst0 + (uint32_t)(c0 >> 6) * (uint32_t)(kPairRA2 * 128) + (uint32_t)row * 128u;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float y[8];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int c = c0 + g * 8 + hh * 4;
              const float4 bb = *reinterpret_cast<const float4*>(prs + 2 * C + c);
              const float4 aa = *reinterpret_cast<const float4*>(prs + 3 * C + c);
              const float4 ii = *reinterpret_cast<const float4*>(prs + 4 * C + c);
              const float4 ww = WS ? *reinterpret_cast<const float4*>(prm + 5 * C + c) : make_float4(1.f, 1.f, 1.f, 1.f);
              y[hh * 4 + 0] = snake_f(fmaf(__uint_as_float(v[g * 8 + hh * 4 + 0]), ww.x, bb.x), aa.x, ii.x);
              y[hh * 4 + 1] = snake_f(fmaf(__uint_as_float(v[g * 8 + hh * 4 + 1]), ww.y, bb.y), aa.y, ii.y);
              y[hh * 4 + 2] = snake_f(fmaf(__uint_as_float(v[g * 8 + hh * 4 + 2]), ww.z, bb.z), aa.z, ii.z);
              y[hh * 4 + 3] = snake_f(fmaf(__uint_as_float(v[g * 8 + hh * 4 + 3]), ww.w, bb.w), aa.w, ii.w);
            }
            uint4 pk = Pack8<ActT>::pack(y);
            if (!valid) pk = make_uint4(0u, 0u, 0u, 0u);
            const uint32_t addr = rowaddr + (uint32_t)(((((c0 & 63) >> 3) + g) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
          }
        }
        tc_fence_before();
        fence_proxy_async();
        if (ew == 0 && lane == 0) trace_ev(a.trace, i, 3);
        mbar_arrive_warp(&d1_empty[b]);
        mbar_arrive_warp(&a2_full[b2]);
      }
      if (do_fin && sr == NS - 1) {
        const int i = it / NS, b = i % NBUF;               // tile index: D2 and its barriers belong to the tile
        const uint32_t u = (uint32_t)(i / NBUF);
        mbar_wait(&d2_full[b], u & 1u);
        tc_fence_after();
        if (warp == 0 && lane == 0) trace_ev(a.trace, it, 4);
        if constexpr (kPreload) {
          const bool has_next = i + NBUF < n_my;
          const ConvTile tnext = has_next ? get_tile(i + NBUF) : tile;
#pragma unroll 1
          for (int blk = grp; blk < NBLK && !(a.dbg & 4); blk += NEPI / 4) {
            float4 pre[8];
            if (has_next) fin_issue(tnext, blk, pre);
            fin_out(tile, blk, b);
            if (has_next) fin_store(tnext, blk, b, pre);
          }
          if (has_next) tmem_st_wait();
          tc_fence_before();
          if (warp == 0 && lane == 0) trace_ev(a.trace, it, 5);
          if (has_next) mbar_arrive_warp(&d2i_full[b]);
        } else {
#pragma unroll
          for (int k = 0; k < BPW; ++k) {
            if (!(a.dbg & 4)) fin_out_res(tile, grp + k * (NEPI / 4), b, pfr[k % PF]);
            pf_issue(i * BPW + k + PF, pfr[k % PF]);
          }
          tc_fence_before();
          if (warp == 0 && lane == 0) trace_ev(a.trace, it, 5);
          mbar_arrive_warp(&d2i_full[b]);
        }
        // combined mode: the staging lives in the A2 tile; nobody may start mid(i+1) before everyone left fin(i)
        if (kCombined) asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (a.mc) cluster_sync_all();              // no CTA leaves while its peer may still multicast into it
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int C, int NBUF, int NA1, int NA2, int W_ST, int NEPI, int NPROD, int NSLAB, int EM, bool kPreload, bool TR, bool WS, typename ActT, int NS = 1>
int launch_pair_ws(const ConvArgs& a, const PairArgs& p, uint32_t idesc, int grid, cudaStream_t st) {
  constexpr int CB = C / 64;
  using PC = PairCfg<C, NBUF, NEPI, NPROD, TR>;
  constexpr int smem = CB * (NA1 * kPairRA1 + NA2 * kPairRA2) * 128 + W_ST * C * 128 + NSLAB * 8192 +
                       (2 * NA1 + 2 * NA2 + 2 * NSLAB + 4 * NBUF + 2 * W_ST) * 8 + 16 + (NS == 1 ? 8 : 5 * NS) * C * 4 +
                       ((PC::kCombined || TR) ? 0 : NEPI * 32 * kStageLd * 4);
  static_assert(smem <= 232448, "shared memory budget exceeded");
  static_assert(!PC::kCombined || NEPI * 32 * kStageLd * 4 <= CB * kPairRA2 * 128, "fin staging must fit the A2 tile");
  static bool configured = false;
  if (!configured) {
    VT_CUDA_OK(cudaFuncSetAttribute(k_pair_tc<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM, kPreload, TR, WS, ActT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (a.mc) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(PC::WARPS * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    VT_CUDA_OK(cudaLaunchKernelEx(&cfg, k_pair_tc<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM, kPreload, TR, WS, ActT, NS>, a, p, idesc));
  } else {
    VT_CUDA_OK(launch_pdl(k_pair_tc<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM, kPreload, TR, WS, ActT, NS>, dim3((unsigned)grid), dim3(PC::WARPS * 32), smem, st, a, p, idesc));
  }
  VT_LAUNCHED();
  return VT_OK;
}

template <int C, int NBUF, int NA1, int NA2, int W_ST, int NEPI, int NPROD, int NSLAB, int EM, bool kPreload, bool TR, typename ActT>
int launch_pair_em(const ConvArgs& a, const PairArgs& p, uint32_t idesc, int grid, cudaStream_t st) {
  // a.wscale / p.wscale1 are null when neither conv of the pair has scaled weight rows
  if (a.wscale && p.wscale1)
    return launch_pair_ws<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM, kPreload, TR, true, ActT>(a, p, idesc, grid, st);
  return launch_pair_ws<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM, kPreload, TR, false, ActT>(a, p, idesc, grid, st);
}

template <int C, int NBUF, int NA1, int NA2, int W_ST, int NEPI, int NPROD, int NSLAB, int EM, typename ActT>
int launch_pair_pre(const ConvArgs& a, const PairArgs& p, uint32_t idesc, int grid, cudaStream_t st) {
  if constexpr (C == 64) {
    if (p.k == 3) return launch_pair_em<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM, false, false, ActT>(a, p, idesc, grid, st);
    return launch_pair_em<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM, true, false, ActT>(a, p, idesc, grid, st);
  } else {
    // VT_PAIR_TR=0: never transposed; =7: only k <= 7 (the k = 11 pairs then run the combined TMEM-preload kernel,
    // which was faster while one thread polled both rings: 0.88 against ~1.0 ms; with separate loader threads the
    // transposed kernel wins for every kernel size - level 1 5.9 -> 5.67 ms)
    static const bool tr = !(getenv("VT_PAIR_TR") && getenv("VT_PAIR_TR")[0] == '0');
    static const bool tr_k7 = getenv("VT_PAIR_TR") && getenv("VT_PAIR_TR")[0] == '7';
    if (tr && (p.k <= 7 || !tr_k7)) {
      // transposed MMA: M = 128 output channels, N = 256 time steps
      const uint32_t idesc_t = (idesc & ~((0x3Fu << 17) | (0x1Fu << 24))) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
      return launch_pair_em<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM, false, true, ActT>(a, p, idesc_t, grid, st);
    }
    return launch_pair_em<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM, true, false, ActT>(a, p, idesc, grid, st);
  }
}

template <int C, int NBUF, int NA1, int NA2, int W_ST, int NEPI, int NPROD, int NSLAB, typename ActT>
int launch_pair_c(const ConvArgs& a, const PairArgs& p, uint32_t idesc, int grid, cudaStream_t st) {
  VT_REQUIRE(a.out && a.res1 && !a.act[1].dst && !a.act[2].dst, "pair_tc: needs an fp32 output and the residual stream");
  const bool oact = a.act[0].dst && a.act[0].kind == ACT_LRELU && a.act_from_out;
  VT_REQUIRE(oact || !a.act[0].dst, "pair_tc: only a leaky-ReLU output copy is supported");
  if (a.res2) {
    VT_REQUIRE(!oact && !a.out_accum && a.out_scale == 1.0f, "pair_tc: unsupported epilogue with two residuals");
    return launch_pair_pre<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM_RES1 | EM_RES2 | EM_OUT, ActT>(a, p, idesc, grid, st);
  }
  if (oact) return launch_pair_pre<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM_RES1 | EM_OUT | EM_ACCUM | EM_OACT, ActT>(a, p, idesc, grid, st);
  if (a.out_accum || a.out_scale != 1.0f)
    return launch_pair_pre<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM_RES1 | EM_OUT | EM_ACCUM, ActT>(a, p, idesc, grid, st);
  return launch_pair_pre<C, NBUF, NA1, NA2, W_ST, NEPI, NPROD, NSLAB, EM_RES1 | EM_OUT, ActT>(a, p, idesc, grid, st);
}

}  // namespace tc

bool pair_tc_supported(const ConvLayer& c1, const ConvLayer& c2) {
  return c1.w_tc && c2.w_tc && c1.cin == c1.cout && c2.cin == c2.cout && c1.cin == c2.cin && (c1.cin == 64 || c1.cin == 128) &&
         c1.k == c2.k && (c1.k & 1) && c2.dil == 1 && (c1.k - 1) * c1.dil <= tc::kPairRA1 - 256 - 6 && c1.k - 1 <= 10 &&
         c1.stride == 1 && c2.stride == 1 && c1.out_mul == 1 && c2.out_mul == 1;
}

int pair_tc_tile_rows(int k) { return 256 - (k - 1); }

namespace {
// debug timeline: VT_TC_TRACE=<conv1 layer name> dumps CTA 0's per-iteration role timestamps to stderr
long long* g_trace = nullptr;
bool trace_begin(ConvArgs& a, const std::string& name, cudaStream_t st) {
  static const char* trace_name = getenv("VT_TC_TRACE");
  if (!(trace_name && name == trace_name)) return false;
  if (!g_trace && cudaMalloc(&g_trace, tc::kTraceTiles * tc::kTraceEvents * 8) != cudaSuccess) return false;
  cudaMemsetAsync(g_trace, 0, tc::kTraceTiles * tc::kTraceEvents * 8, st);
  a.trace = g_trace;
  return true;
}
int trace_dump(const ConvArgs& a, const std::string& name, int k, int dil, int grid, cudaStream_t st) {
  std::vector<long long> h(tc::kTraceTiles * tc::kTraceEvents);
  VT_CUDA_OK(cudaStreamSynchronize(st));
  VT_CUDA_OK(cudaMemcpy(h.data(), g_trace, h.size() * 8, cudaMemcpyDeviceToHost));
  long long t0 = 0;
  for (size_t q = 0; q < h.size(); ++q)
    if ((int)(q % tc::kTraceEvents) < 10 && h[q] && (!t0 || h[q] < t0)) t0 = h[q];
  fprintf(stderr, "[vt trace] pair %s k=%d dil=%d tiles=%d grid=%d (cycles; PROD start end | MID start end | FIN start end | "
          "C1 start issued | C2 start issued)\n", name.c_str(), k, dil, a.n_tiles, grid);
  for (int it = 0; it < tc::kTraceTiles; ++it) {
    if (!h[it * tc::kTraceEvents + 0]) break;
    fprintf(stderr, "[vt trace] %2d", it);
    for (int e = 0; e < 10; ++e) fprintf(stderr, " %7lld", h[it * tc::kTraceEvents + e] ? h[it * tc::kTraceEvents + e] - t0 : -1);
    fprintf(stderr, "  w_wait c1=%lld c2=%lld x_wait=%lld", h[it * tc::kTraceEvents + 10], h[it * tc::kTraceEvents + 11], h[it * tc::kTraceEvents + 12]);
    fprintf(stderr, "\n");
  }
  return VT_OK;
}
}  // namespace

// `a` describes the fin epilogue (bias = conv2 bias, res1 = the pair's input stream, out, tiles of
// pair_tc_tile_rows(k) output steps); alpha1 / alpha2 are the Snake parameters before conv1 / conv2.
int launch_pair_tc(const ConvArgs& a_in, const ConvLayer& c1, const ConvLayer& c2, const float* alpha1, const float* alpha2,
                   int act_elem, cudaStream_t st) {
  VT_REQUIRE(pair_tc_supported(c1, c2) && (act_elem == ELEM_F16 || act_elem == ELEM_BF16), "pair_tc: layers %s / %s cannot be fused",
             c1.name.c_str(), c2.name.c_str());
  if (a_in.n_tiles == 0) return VT_OK;
  ConvArgs a = a_in;
  static const int dbg = getenv("VT_TC_DBG") ? atoi(getenv("VT_TC_DBG")) : 0;
  a.dbg = dbg;
  a.bias = c2.bias;
  a.cout = c2.cout; a.phase_c = c2.cout; a.out_mul = 1; a.out_shift = 0; a.dup_row2 = 0;
  PairArgs p{};
  p.x_in = a.res1;
  p.alpha1 = alpha1; p.alpha2 = alpha2; p.bias1 = c1.bias;
  // scaled weight rows in either conv -> the WS instance (an unscaled partner carries a vector of ones)
  const bool ws = c1.scaled || c2.scaled;
  p.wscale1 = ws ? c1.wscale : nullptr;
  a.wscale = ws ? c2.wscale : nullptr;
  p.w1 = reinterpret_cast<const uint8_t*>(c1.w_tc);
  p.w2 = reinterpret_cast<const uint8_t*>(c2.w_tc);
  p.k = c1.k; p.dil = c1.dil;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    VT_CUDA_OK(cudaGetDevice(&dev));
    VT_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int C = c1.cin;
  // VT_PAIR_MC=1: the two CTAs of a cluster share one weight stream (each fetches half of every ring slot and
  // multicasts it).  Verified, but measured equal to per-CTA streams (29.85 against 29.84-30.13 ms per forward): the
  // cost of the weight stream is on the SM side (shared-memory writes), not in L2 reads - off by default.
  static const bool use_mc = getenv("VT_PAIR_MC") && getenv("VT_PAIR_MC")[0] == '1';
  a.mc = use_mc ? 1 : 0;
  const int n_even = (a.n_tiles + 1) & ~1;
  const int grid = a.mc ? (n_even < (sm_count & ~1) ? n_even : (sm_count & ~1)) : (a.n_tiles < sm_count ? a.n_tiles : sm_count);
  const bool tracing = trace_begin(a, c1.name, st);
  const uint32_t fmt = act_elem == ELEM_F16 ? 0u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(C >> 3) << 17) | ((128u >> 4) << 24);
  int rc;
  // experiment switch: C = 128 with a 2-slot weight ring and a 5-slab x ring (same shared memory)
  static const bool ring5 = getenv("VT_P128_RING") && getenv("VT_P128_RING")[0] == '5';
  static const bool nprod4 = getenv("VT_P64_NPROD") && getenv("VT_P64_NPROD")[0] == '4';   // experiment: 4 producer warps at C = 64
  if (act_elem == ELEM_F16 && C == 128 && ring5)
    rc = tc::launch_pair_c<128, 1, 1, 1, 2, 8, 4, 5, __half>(a, p, idesc, grid, st);
  else if (act_elem == ELEM_F16 && C == 64 && nprod4)
    rc = tc::launch_pair_c<64, 2, 2, 1, 4, 8, 4, 5, __half>(a, p, idesc, grid, st);
  else if (act_elem == ELEM_F16)
    rc = C == 64 ? tc::launch_pair_c<64, 2, 2, 1, 4, 8, 2, 5, __half>(a, p, idesc, grid, st)
                 : tc::launch_pair_c<128, 1, 1, 1, 3, 8, 4, 3, __half>(a, p, idesc, grid, st);
  else
    rc = C == 64 ? tc::launch_pair_c<64, 2, 2, 1, 4, 8, 2, 5, __nv_bfloat16>(a, p, idesc, grid, st)
                 : tc::launch_pair_c<128, 1, 1, 1, 3, 8, 4, 3, __nv_bfloat16>(a, p, idesc, grid, st);
  if (tracing && rc == VT_OK) return trace_dump(a, c1.name, c1.k, c1.dil, grid, st);
  return rc;
}

// Mean-fused launch of the LAST pairs of the three ResBlocks of a stage (k_pair_tc, NS = 3):
//   out = (1/3) sum_r [ x_r + conv2_r(Snake(conv1_r(Snake(x_r)))) ]   (+ the next layer's leaky-ReLU operand copy)
// `a` describes the fin epilogue as for launch_pair_tc (out, out_scale = 1/3, act[0] = the leaky-ReLU copy) on the tiles of
// the LARGEST kernel size (pair_tc_tile_rows(k_max) output steps: every sub is evaluated on the same 256 conv1 rows, a
// smaller kernel only needs less halo).
bool pair3_tc_supported(const ConvLayer* const c1[3], const ConvLayer* const c2[3], int act_elem) {
  // VT_PAIR3=0 disables the mean-fused launch everywhere, VT_PAIR3=64 keeps it to C = 64.  At C = 128 (D2 single buffered:
  // TMEM is full) conv2 of the next tile's first sub waits for the whole fin pass, which loads three residual streams:
  // with the general fin pass (clamped indices, 112 bytes of spills at the 80-register cap) level 1 went 7.43 -> 7.89 ms;
  // with the full-block fast path the instance no longer spills and the launch wins there too (7.21 -> 7.13 ms).
  static const char* sw = getenv("VT_PAIR3");
  const bool off = sw && sw[0] == '0', only64 = sw && sw[0] == '6';
  if (off || act_elem != ELEM_F16 || (c1[0]->cin != 64 && only64)) return false;
  for (int r = 0; r < 3; ++r)
    if (!pair_tc_supported(*c1[r], *c2[r]) || c1[r]->scaled || c2[r]->scaled || c1[r]->cin != c1[0]->cin) return false;
  return true;
}

int launch_pair3_tc(const ConvArgs& a_in, const ConvLayer* const c1[3], const ConvLayer* const c2[3], const float* const alpha1[3],
                    const float* const alpha2[3], const float* const x_in[3], int act_elem, cudaStream_t st) {
  VT_REQUIRE(pair3_tc_supported(c1, c2, act_elem), "pair3_tc: layers %s .. cannot be mean-fused", c1[0]->name.c_str());
  if (a_in.n_tiles == 0) return VT_OK;
  ConvArgs a = a_in;
  static const int dbg = getenv("VT_TC_DBG") ? atoi(getenv("VT_TC_DBG")) : 0;
  a.dbg = dbg;
  VT_REQUIRE(a.out && !a.res2 && !a.out_accum && a.act[0].dst && a.act[0].kind == ACT_LRELU && a.act_from_out && !a.act[1].dst,
             "pair3_tc: needs an fp32 output and the leaky-ReLU operand copy, nothing else");
  a.bias = c2[0]->bias; a.res1 = x_in[0]; a.wscale = nullptr; a.mc = 0; a.trace = nullptr;
  const bool tracing = trace_begin(a, c1[0]->name, st);
  a.cout = c2[0]->cout; a.phase_c = c2[0]->cout; a.out_mul = 1; a.out_shift = 0; a.dup_row2 = 0;
  PairArgs p{};
  p.x_in = x_in[0]; p.alpha1 = alpha1[0]; p.alpha2 = alpha2[0]; p.bias1 = c1[0]->bias;
  p.w1 = reinterpret_cast<const uint8_t*>(c1[0]->w_tc); p.w2 = reinterpret_cast<const uint8_t*>(c2[0]->w_tc);
  p.k = c1[0]->k; p.dil = c1[0]->dil;
  p.nsub = 3;
  for (int r = 1; r < 3; ++r) {
    PairArgs::Sub& m = p.more[r - 1];
    m.x_in = x_in[r]; m.alpha1 = alpha1[r]; m.alpha2 = alpha2[r]; m.bias1 = c1[r]->bias; m.bias2 = c2[r]->bias;
    m.w1 = reinterpret_cast<const uint8_t*>(c1[r]->w_tc); m.w2 = reinterpret_cast<const uint8_t*>(c2[r]->w_tc);
    m.k = c1[r]->k; m.dil = c1[r]->dil;
  }
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    VT_CUDA_OK(cudaGetDevice(&dev));
    VT_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int C = c1[0]->cin;
  const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
  constexpr int EM = tc::EM_RES1 | tc::EM_OUT | tc::EM_OACT;
  int rc;
  if (C == 64) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
    rc = tc::launch_pair_ws<64, 2, 2, 1, 4, 8, 2, 5, EM, true, false, false, __half, 3>(a, p, idesc, grid, st);
  } else {
    // transposed MMA: M = 128 output channels, N = 256 time steps
    const uint32_t idesc_t = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    rc = tc::launch_pair_ws<128, 1, 1, 1, 3, 8, 4, 3, EM, false, true, false, __half, 3>(a, p, idesc_t, grid, st);
  }
  if (tracing && rc == VT_OK) return trace_dump(a, c1[0]->name, c1[0]->k, c1[0]->dil, grid, st);
  return rc;
}

}  // namespace vt
