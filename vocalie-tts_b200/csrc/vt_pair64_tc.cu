// Fused ResBlock pair at C = 64 with TAP-PAIRED transposed MMAs (upstream hifigan.py ResBlock.forward, one
// iteration of its dilation loop; same contract as vt_pair_tc.cu).
//
// Why another formulation: with activations as the M operand an MMA at C = 64 is 128 x 64 x 16 - the tensor core
// needs 32 cycles for it, the 4 KB activation fetch from shared memory 68 (measured), so the k = 11 pairs of the
// largest level run the tensor pipe at a third of its rate and are bound by exactly that.  Transposed (weights as
// the M operand, 256 time steps as N) M would be 64: half rate again.  Here TWO TAPS are stacked in M:
//     64 rows of the A operand = W[tap e]   ("E" half)
//     64 rows of the A operand = W[tap e+1] ("L" half)
// and both halves multiply the SAME row-shifted activation window, so one 128 x 256 x 16 MMA does the work of four
// 128 x 64 x 16 ones.  The price: the two halves accumulate contributions to output steps that differ by one tap,
//     out[t] = E[t] + L[t + s],  s = dil (conv1) or 1 (conv2);  pairs (0,1), (2,3), .., (k-1, zero); window shift 2p*s
// i.e. a COLUMN shift (free: a TMEM address) between the two halves.  The rows of the A operand are permuted so
// that the two halves of a channel sit 16 lanes apart in the SAME lane quarter (rows 32q..32q+15 = E of channels
// 16q.., rows 32q+16..32q+31 = L of the same channels): the sum is one warp shuffle, no shared memory, no barrier.
// The last 8-column block of a tile would read past the accumulator: its address is clamped, so columns >= 248 are
// garbage - conv1 columns < 248 feed conv2 columns < 249 - k, the output steps of a tile.
//
// Per CTA tile (256 window columns = 249 - k output steps):
//   2 loader threads  weight ring (16 KB slots: one tap pair each) and x ring (8 KB slabs of fp32 rows), bulk copies
//   producers (4)   x ring --Snake1--> fp16 A1 tile (SWIZZLE_128B, rows = time)
//   MMA warp        step s: conv1(s) -> D1, then conv2(s-1) -> D2   (N = 256 columns each: TMEM is full)
//   epilogue (16)   step s: mid(s): D1 + b1 --Snake2--> A2[s & 1] (zero outside the sequence);
//                           fin(s-1): D2 + b2 + x (+ second residual / running mean) -> fp32 stream (+ leaky-ReLU copy)
// mid(s) runs while the tensor pipe works on conv2(s-1), fin(s-1) while it works on conv1(s+1).
//
// Measured (B200, 64 x 500 frames, k = 11): conv1 + conv2 of a tile issue in 7 k cycles (16.5 k on the activation-
// major kernel); the tile period is 11-12 k cycles and is set by the CUDA-core roles - a producer warp is alone on its
// scheduler and converts a 32-row slab in ~600 cycles whatever else runs (latency-bound instruction stream, not the SFU:
// moving half of the sines to the FMA pipe made it slower), mid + fin take 8 k cycles on the 16 epilogue warps.  Per
// launch (ncu): 0.86-0.90 ms against 1.01 ms for the plain pairs.  The LAST pair of a ResBlock (second residual,
// running mean, operand copy) is slower here (1.2-1.4 against 1.05-1.07 ms): its extra streams are loaded with dependent
// arithmetic in issue_x(), whose latency lands on the mid -> conv1 chain; vt_hift.cu keeps those on vt_pair_tc.cu.
#include "vt_tc.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace vt {

namespace tc {

constexpr int kP64RA1 = 312;                 // A1 rows: 256 + (k-1)*dil <= 306
constexpr int kP64RA2 = 272;                 // A2 rows: 256 + (k-1) <= 266
constexpr int kP64NPROD = 4, kP64NEPI = 16;
constexpr int kP64A1Bytes = kP64RA1 * 128, kP64A2Bytes = kP64RA2 * 128, kP64WBytes = 16384, kP64SlabBytes = 8192;
constexpr int kP64SlabRows = kP64SlabBytes / 256;
constexpr int kP64W_MMA = kP64NEPI, kP64W_LD = kP64NEPI + 1, kP64W_LX = kP64NEPI + 2, kP64W_AP = kP64NEPI + 3;
constexpr int kP64Warps = kP64NEPI + 3 + kP64NPROD;     // 23: the register file is allocated for 24 anyway
// NA1 = A1 buffers: 2 leaves room for 3 weight + 3 x ring slots, 1 for 4 + 5
template <int NA1> struct P64Cfg {
  static constexpr int W_ST = NA1 == 2 ? 3 : 4, NSLAB = NA1 == 2 ? 3 : 5;
  static constexpr int kBars = 4 + 2 * NSLAB + 4 + 4 + 2 * W_ST;
  static constexpr int kSmem = NA1 * kP64A1Bytes + 2 * kP64A2Bytes + W_ST * kP64WBytes + NSLAB * kP64SlabBytes + kBars * 8 + 16 + 5 * 64 * 4;
  static_assert(kSmem <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ void p64_tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
template <typename T> __device__ __forceinline__ unsigned short p64_op_bits(float v);
template <> __device__ __forceinline__ unsigned short p64_op_bits<__half>(float v) { return __half_as_ushort(__float2half_rn(v)); }
template <> __device__ __forceinline__ unsigned short p64_op_bits<__nv_bfloat16>(float v) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}

template <int EM, int NA1, typename ActT>
__global__ void __launch_bounds__(kP64Warps * 32, 1) k_pair64_tc(const ConvArgs a, const PairArgs p, const uint32_t idesc) {
  constexpr int C = 64, NSLAB = P64Cfg<NA1>::NSLAB, W_ST = P64Cfg<NA1>::W_ST, NEPI = kP64NEPI, kProdT = kP64NPROD * 32;
  constexpr int kP64NA1 = NA1;
  static_assert(kSwz, "the fused pair kernel assumes the SWIZZLE_128B operand layout");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA1 = smem;
  uint8_t* sA2 = sA1 + kP64NA1 * kP64A1Bytes;
  uint8_t* sW = sA2 + 2 * kP64A2Bytes;
  uint8_t* sX = sW + W_ST * kP64WBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + NSLAB * kP64SlabBytes);
  uint64_t* a1_full = bars;
  uint64_t* a1_empty = a1_full + 2;      // [2]
  uint64_t* x_full = a1_empty + 2;
  uint64_t* x_empty = x_full + NSLAB;
  uint64_t* a2_full = x_empty + NSLAB;       // [2]
  uint64_t* a2_empty = a2_full + 2;          // [2]
  uint64_t* d1_full = a2_empty + 2;
  uint64_t* d1_empty = d1_full + 1;
  uint64_t* d2_full = d1_empty + 1;
  uint64_t* d2_empty = d2_full + 1;
  uint64_t* w_full = d2_empty + 1;
  uint64_t* w_empty = w_full + W_ST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + W_ST);
  float* prm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tmem_slot) + 16);   // al1 | ia1 | b1 | al2 | ia2

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    // consumer-side barriers count WARPS: one elected arrive per warp after __syncwarp (32 arrives of a warp on one
    // barrier serialise in the shared-memory atomic unit)
    for (int i = 0; i < 2; ++i) { mbar_init(&a1_full[i], kP64NPROD); mbar_init(&a1_empty[i], 1); }
    for (int i = 0; i < NSLAB; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], kP64NPROD); }
    for (int i = 0; i < 2; ++i) { mbar_init(&a2_full[i], NEPI); mbar_init(&a2_empty[i], 1); }
    mbar_init(d1_full, 1); mbar_init(d1_empty, NEPI);
    mbar_init(d2_full, 1); mbar_init(d2_empty, NEPI);
    for (int i = 0; i < W_ST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    fence_barrier_init();
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float a1 = p.alpha1[c], a2 = p.alpha2[c];
    prm[c] = a1; prm[C + c] = __fdividef(1.0f, a1 + 1e-9f); prm[2 * C + c] = p.bias1[c];
    prm[3 * C + c] = a2; prm[4 * C + c] = __fdividef(1.0f, a2 + 1e-9f);
  }
  if (warp == kP64W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();                             // dependents may be scheduled (they take an SM when its CTA of this grid exits)
  if (warp != kP64W_LD) pdl_wait();                // everything but the (static) weight stream waits for the previous kernel

  const int n_tiles = a.n_tiles;
  const int n_my = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int H2 = (p.k - 1) / 2, H1 = (p.k - 1) * p.dil / 2;
  const int npairs = (p.k + 1) / 2;
  const int R1 = 256 + 2 * H1;               // A1 rows the MMAs touch

  if (warp >= kP64W_AP) {
    // ---------------- producers: fp32 slabs of the x ring -> Snake1 -> fp16 A1 rows.  A thread owns the channels
    // [4*ch, 4*ch+4) and [32 + 4*ch, +4) of every row it touches, so its Snake parameters stay in registers.
    constexpr int QPR = C / 8, RPP = kProdT / QPR, TPS = kP64SlabRows / RPP;
    static_assert(kP64SlabRows % RPP == 0, "slab rows must divide among the producer threads");
    const int pt = threadIdx.x - kP64W_AP * 32;
    const int ch = pt % QPR, r_raw = pt / QPR;
    // bits 0 and 2 swapped: the two rows of a HALF warp (the unit an 8-byte store is served in) lie 4 rows apart, on both
    // sides of the swizzle pattern's (row & 4) - their 16 stores then cover all 32 banks once (see vt_pair_tc.cu)
#ifdef VT_OLD_ROWPERM
    const int r_in = (r_raw & ~6) | ((r_raw & 2) << 1) | ((r_raw & 4) >> 1);
#else
    const int r_in = (r_raw & ~5) | ((r_raw & 1) << 2) | ((r_raw & 4) >> 2);
#endif
    const int cA = 4 * ch, cB = C / 2 + 4 * ch;
    float alA[4], iaA[4], alB[4], iaB[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      alA[e] = prm[cA + e]; iaA[e] = prm[C + cA + e];
      alB[e] = prm[cB + e]; iaB[e] = prm[C + cB + e];
    }
    const int n_slab = (R1 + kP64SlabRows - 1) / kP64SlabRows;
    const uint32_t chkA = (uint32_t)(cA >> 3), chkB = (uint32_t)(cB >> 3);
    const uint32_t halfA = (uint32_t)((cA & 7) >> 2) * 8u, halfB = (uint32_t)((cB & 7) >> 2) * 8u;
    // One producer warp per scheduler: nothing hides its shared-memory latency but its own instruction stream, so
    // the reads of slab g+1 (also across tiles: the x ring does not depend on the A1 buffers) are issued before slab
    // g is converted.
    uint32_t xs = 0, xph = 0;
    float4 ca[TPS], cb[TPS], na[TPS], nb[TPS];
    uint32_t slot_c = 0, slot_n = 0;
    long long x_wait = 0;
    auto load_slab = [&](float4 (&va)[TPS], float4 (&vb)[TPS], uint32_t& slot) {
      if (a.trace) {
        const long long tw = clock64();
        mbar_wait(&x_full[xs], xph);
        x_wait += clock64() - tw;
      } else {
        mbar_wait(&x_full[xs], xph);
      }
      const uint32_t xsrc = smem_u32(sX + xs * kP64SlabBytes);
#pragma unroll
      for (int t = 0; t < TPS; ++t) {
        const int rr = r_in + t * RPP;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(va[t].x), "=f"(va[t].y), "=f"(va[t].z), "=f"(va[t].w)
                     : "r"(xsrc + (uint32_t)(rr * C * 4 + cA * 4)));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vb[t].x), "=f"(vb[t].y), "=f"(vb[t].z), "=f"(vb[t].w)
                     : "r"(xsrc + (uint32_t)(rr * C * 4 + cB * 4)));
      }
      slot = xs;
      if (++xs == (uint32_t)NSLAB) { xs = 0; xph ^= 1u; }
    };
    const int total = n_my * n_slab;
    int ti = 0, sl = 0;
    // one slab of the stream: `cur` is converted while `nxt` is in flight (two register sets, swapped by the caller)
    auto step = [&](float4 (&va)[TPS], float4 (&vb)[TPS], uint32_t& slot_cur, float4 (&wa)[TPS], float4 (&wb)[TPS], uint32_t& slot_nxt,
                    int g) {
      const uint32_t a1 = smem_u32(sA1) + (uint32_t)(ti % kP64NA1) * (uint32_t)kP64A1Bytes;
      if (sl == 0) {
        mbar_wait(&a1_empty[ti % kP64NA1], ((uint32_t)(ti / kP64NA1) & 1u) ^ 1u);
        if (pt == 0) trace_ev(a.trace, ti, 0);
      }
      if (g + 1 < total) load_slab(wa, wb, slot_nxt);
#pragma unroll
      for (int t = 0; t < TPS; ++t) {
        const int r = sl * kP64SlabRows + r_in + t * RPP;
        if (r < R1) {
          float ya[4] = {snake_f(va[t].x, alA[0], iaA[0]), snake_f(va[t].y, alA[1], iaA[1]), snake_f(va[t].z, alA[2], iaA[2]),
                         snake_f(va[t].w, alA[3], iaA[3])};
          float yb[4] = {snake_f(vb[t].x, alB[0], iaB[0]), snake_f(vb[t].y, alB[1], iaB[1]), snake_f(vb[t].z, alB[2], iaB[2]),
                         snake_f(vb[t].w, alB[3], iaB[3])};
          const uint2 pa = Pack4<ActT>::pack(ya), pb = Pack4<ActT>::pack(yb);
          const uint32_t rowb = a1 + (uint32_t)r * 128u;
          const uint32_t swz = (uint32_t)(r & 7);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowb + ((chkA ^ swz) << 4) + halfA), "r"(pa.x), "r"(pa.y) : "memory");
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowb + ((chkB ^ swz) << 4) + halfB), "r"(pb.x), "r"(pb.y) : "memory");
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[slot_cur]);   // the slab is in registers and converted: the loader may refill it
      if (++sl == n_slab) {
        fence_proxy_async();
        if (pt == 0) trace_ev(a.trace, ti, 1);
        if (pt == 0 && a.trace && blockIdx.x == 0 && ti < kTraceTiles) { a.trace[ti * kTraceEvents + 12] = x_wait; x_wait = 0; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&a1_full[ti % kP64NA1]);
        sl = 0;
        ++ti;
      }
    };
    if (total > 0) load_slab(ca, cb, slot_c);
    for (int g = 0; g < total; g += 2) {
      step(ca, cb, slot_c, na, nb, slot_n, g);
      if (g + 1 < total) step(na, nb, slot_n, ca, cb, slot_c, g + 1);
    }
  } else if (warp == kP64W_LD) {
    // ---------------- weight loader: one tap pair (16 KB) per ring slot, in the MMA warp's order - step s: conv1(s)
    // if s < n_my, then conv2(s-1) if s >= 1
    if (lane == 0) {
      uint32_t ws = 0, wph = 0;
      for (int s = 0; s <= n_my; ++s)
        for (int pass = 0; pass < 2; ++pass) {
          if (pass == 0 ? s >= n_my : s < 1) continue;
          const uint8_t* wsrc = pass == 0 ? p.w1 : p.w2;         // pack_pair64 image: one 16 KB slab per tap pair
          for (int pp = 0; pp < npairs; ++pp) {
            mbar_wait(&w_empty[ws], wph ^ 1u);
            if (a.dbg & 1) mbar_arrive(&w_full[ws]);
            else {
              mbar_arrive_expect_tx(&w_full[ws], kP64WBytes);
              bulk_g2s(sW + ws * kP64WBytes, wsrc + (size_t)pp * kP64WBytes, kP64WBytes, &w_full[ws]);
            }
            if (++ws == (uint32_t)W_ST) { ws = 0; wph ^= 1u; }
          }
        }
    }
  } else if (warp == kP64W_LX) {
    // ---------------- x loader: the tiles' fp32 rows (with halo), slab by slab (whole rows are contiguous in HBM ->
    // 1-D bulk copies).  Its own thread: a slot is refilled the moment the producers release it.
    if (lane == 0) {
      uint32_t xs = 0, xph = 0;
      ConvTile tl = n_my > 0 ? a.tiles[blockIdx.x] : ConvTile{};
      for (int xi = 0; xi < n_my; ++xi) {
        const float* xsrc = p.x_in + (tl.in_row0 + tl.q0 - H1 - H2) * (long long)C;
        if (xi + 1 < n_my) {
          tl = a.tiles[blockIdx.x + (xi + 1) * gridDim.x];      // next tile's entry: off the critical path
          // The ring holds 24-40 KB and a bulk copy from HBM takes ~2 us under load; the next tile's rows are pulled
          // into L2 a whole tile period ahead (measured: tile period 11.3 k -> 11.1 k cycles - the ring is not what
          // bounds the producers).
          if (!(a.dbg & 512)) {
            const float* nsrc = p.x_in + (tl.in_row0 + tl.q0 - H1 - H2) * (long long)C;
            for (int xr = 0; xr < R1; xr += kP64SlabRows) {
              const int rows = R1 - xr < kP64SlabRows ? R1 - xr : kP64SlabRows;
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nsrc + (long long)xr * C), "r"((uint32_t)(rows * C * 4)) : "memory");
            }
          }
        }
        for (int xr = 0; xr < R1; xr += kP64SlabRows) {
          const int rows = R1 - xr < kP64SlabRows ? R1 - xr : kP64SlabRows;
          mbar_wait(&x_empty[xs], xph ^ 1u);
          if (a.dbg & 2) mbar_arrive(&x_full[xs]);
          else {
            mbar_arrive_expect_tx(&x_full[xs], (uint32_t)(rows * C * 4));
            bulk_g2s(sX + xs * kP64SlabBytes, xsrc + (long long)xr * C, (uint32_t)(rows * C * 4), &x_full[xs]);
          }
          if (++xs == (uint32_t)NSLAB) { xs = 0; xph ^= 1u; }
        }
      }
    }
  } else if (warp == kP64W_MMA) {
    // ---------------- MMA issuer
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a1_lo = ((smem_u32(sA1) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t a2_lo0 = ((smem_u32(sA2) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t w_lo0 = ((smem_u32(sW) >> 4) & 0x3FFFu) | (1u << 16);
    const bool mma_on = !(a.dbg & 16);
    uint32_t ws = 0, wph = 0;
    for (int s = 0; s <= n_my; ++s)
      for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0 ? s >= n_my : s < 1) continue;
        const int i = pass == 0 ? s : s - 1;
        const uint32_t ui = (uint32_t)i;
        if (pass == 0) {
          mbar_wait(d1_empty, (ui & 1u) ^ 1u);                 // mid(i-1) has drained D1
          mbar_wait(&a1_full[i % kP64NA1], (uint32_t)(i / kP64NA1) & 1u);
        } else {
          mbar_wait(d2_empty, (ui & 1u) ^ 1u);                 // fin(i-1) has drained D2
          mbar_wait(&a2_full[i & 1], (ui >> 1) & 1u);
        }
        tc_fence_after();
        if (lane == 0) trace_ev(a.trace, i, 6 + 2 * pass);
        const uint32_t d0 = tmem_base + (uint32_t)(pass * 256);
        const uint32_t b_tile = pass == 0 ? a1_lo + (uint32_t)(i % kP64NA1) * (uint32_t)(kP64A1Bytes >> 4) : a2_lo0 + (uint32_t)(i & 1) * (uint32_t)(kP64A2Bytes >> 4);
        const uint32_t pair16 = (uint32_t)(pass == 0 ? 2 * p.dil : 2) * 8u;   // one tap pair = 2 taps of `dil` rows of 128 B
        long long w_wait = 0;
        for (int pp = 0; pp < npairs; ++pp) {
          if (a.trace) {
            const long long tw = clock64();
            mbar_wait(&w_full[ws], wph);
            w_wait += clock64() - tw;
          } else {
            mbar_wait(&w_full[ws], wph);
          }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t w_lo = w_lo0 + ws * (uint32_t)(kP64WBytes >> 4);
            const uint32_t b_chunk = b_tile + (uint32_t)pp * pair16;
            if (mma_on) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_f16_lh(d0, w_lo + (uint32_t)(ks * 2), b_chunk + (uint32_t)(ks * 2), kDescHi, idesc, (pp | ks) ? 1u : 0u);
            }
            umma_commit(&w_empty[ws]);
          }
          __syncwarp();
          if (++ws == (uint32_t)W_ST) { ws = 0; wph ^= 1u; }
        }
        if (elect_one()) {
          umma_commit(pass == 0 ? &a1_empty[i % kP64NA1] : &a2_empty[i & 1]);
          umma_commit(pass == 0 ? d1_full : d2_full);
        }
        __syncwarp();
        if (lane == 0) trace_ev(a.trace, i, 7 + 2 * pass);
        if (lane == 0 && a.trace && blockIdx.x == 0 && i < kTraceTiles) a.trace[i * kTraceEvents + 10 + pass] = w_wait;
      }
  } else {
    // ---------------- epilogue warps.  The A operand rows are permuted (pack_pair64) so that every lane quarter holds
    // BOTH halves of 16 channels: lanes 0..15 of quarter q = E of channels 16q.., lanes 16..31 = L of the same
    // channels.  out[t] = E[t] + L[t + s] is then a shuffle between lanes l and l ^ 16 of ONE warp: per 8-column
    // block the low lanes finalise columns 0..3, the high lanes columns 4..7.  Warp = (quarter, 64-column range tq).
    constexpr bool kRes2 = (EM & EM_RES2) != 0, kAccum = (EM & EM_ACCUM) != 0, kOact = (EM & EM_OACT) != 0;
    const int quarter = warp & 3, tq = warp >> 2;
    const bool hi = lane >= 16;
    const int c = quarter * 16 + (lane & 15);                             // this thread's channel
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const int col_lo = 64 * tq + (hi ? 4 : 0);                            // first column this thread finalises
    const float b1 = prm[2 * C + c], al2 = prm[3 * C + c], ia2 = prm[4 * C + c];
    const float b2 = a.bias[c], ws1 = p.wscale1[c], ws2 = a.wscale[c];   // inverse power-of-two weight-row scales
    const bool accum = kAccum && a.out_accum;
    const float inv = 1.0f / a.out_scale;
    const uint32_t a2_off = (uint32_t)((c & 7) * 2);
    const uint32_t a2_chunk = (uint32_t)(c >> 3);

    // accumulator columns [t0, t0+8) of the E lanes and [t0+s, t0+s+8) of the L lanes (address clamped: see the header)
    auto ld_block = [&](uint32_t dbase, int t0, int s, uint32_t (&v1)[8], uint32_t (&v2)[8]) {
      p64_tmem_ld8(dbase + lane_sel + (uint32_t)t0, v1);
      p64_tmem_ld8(dbase + lane_sel + (uint32_t)(t0 + s < 248 ? t0 + s : 248), v2);
    };
    // the 4 sums this thread finalises: low lanes columns t0..t0+3, high lanes t0+4..t0+7
    auto combine = [&](const uint32_t (&v1)[8], const uint32_t (&v2)[8], float (&sum)[4]) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float send = __uint_as_float(hi ? v2[jj] : v1[jj + 4]);     // what lane ^ 16 needs from this lane
        const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        sum[jj] = __uint_as_float(hi ? v2[jj + 4] : v1[jj]) + recv;
      }
    };

    // The instruction issue slots of the four schedulers are a real budget in this kernel (16 epilogue warps + 4
    // producers), so the common case - every column of this warp's range inside the sequence and inside the tile - runs
    // without per-element predicates and with compile-time address offsets.
    // fin: the residual terms of this thread's 32 outputs of a tile are requested a whole step ahead (load latency
    // under the weight stream is microseconds)
    float xr[32];
    auto tile_base = [&](int i, int& n) {
      const ConvTile tl = a.tiles[blockIdx.x + i * gridDim.x];
      n = tl.n;
      return (tl.out_row0 + tl.q0 + col_lo) * (long long)C + c;           // element (first column of this thread, channel)
    };
    auto issue_x = [&](long long base, int n) {
      if (64 * tq + 64 <= n) {
        const float* xp = a.res1 + base;
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) xr[4 * b + jj] = (a.dbg & 32) ? 0.0f : __ldg(xp + (8 * b + jj) * C);
        if constexpr (kRes2) {
          const float* rp = a.res2 + base;
#pragma unroll
          for (int q = 0; q < 32; ++q) xr[q] += __ldg(rp + (8 * (q >> 2) + (q & 3)) * C);
        }
        if (accum) {
          const float* pp = a.out + base;
#pragma unroll
          for (int q = 0; q < 32; ++q) xr[q] = fmaf(__ldg(pp + (8 * (q >> 2) + (q & 3)) * C), inv, xr[q]);
        }
      } else {
        // ragged tile: clamped rows (a branch per load would serialise them)
        const int last = n - 1 - col_lo;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int rel = 8 * (q >> 2) + (q & 3);
          const long long off = base + (long long)(rel < last ? rel : last) * C;
          float x = (a.dbg & 32) ? 0.0f : __ldg(a.res1 + off);
          if constexpr (kRes2) x += __ldg(a.res2 + off);
          if (accum) x = fmaf(__ldg(a.out + off), inv, x);
          xr[q] = x;
        }
      }
    };
    auto ewait = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };
    int n_next = 1;
    long long base_next = n_my > 0 ? tile_base(0, n_next) : 0;
    if (n_my > 0) issue_x(base_next, n_next);
    // A2 byte offsets of this thread's 4 rows inside an 8-row group (row & 7 = (hi ? 4 : 0) + jj: the swizzle phase)
    uint32_t a2_rel[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const uint32_t r7 = (uint32_t)((hi ? 4 : 0) + jj);
      a2_rel[jj] = (uint32_t)(64 * tq) * 128u + r7 * 128u + ((a2_chunk ^ r7) << 4) + a2_off;
    }

    for (int s = 0; s <= n_my; ++s) {
      if (s < n_my) {
        // ---- mid(s): conv1 out[t] = E[t] + L[t + dil] -> + b1 -> Snake2 -> A2 row t
        const ConvTile tile = a.tiles[blockIdx.x + s * gridDim.x];
        const int p0 = tile.q0 - H2 + 64 * tq;                               // sequence position of this warp's first column
        const bool inside = p0 >= 0 && p0 + 64 <= tile.out_len;
        ewait(d1_full, (uint32_t)s & 1u);
        ewait(&a2_empty[s & 1], (((uint32_t)s >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (warp == 0 && lane == 0) trace_ev(a.trace, s, 2);
        uint8_t* dstp = sA2 + (s & 1) * kP64A2Bytes;
        if (!(a.dbg & 4)) {
          uint32_t v1[2][8], v2[2][8];
          ld_block(tmem_base, 64 * tq, p.dil, v1[0], v2[0]);
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            tmem_ld_wait();
            if (b < 7) ld_block(tmem_base, 64 * tq + 8 * (b + 1), p.dil, v1[(b + 1) & 1], v2[(b + 1) & 1]);
            float sum[4];
            combine(v1[b & 1], v2[b & 1], sum);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float y = snake_f(fmaf(sum[jj], ws1, b1), al2, ia2);
              if (!inside) {
                const int pseq = p0 + (hi ? 4 : 0) + 8 * b + jj;
                if (pseq < 0 || pseq >= tile.out_len) y = 0.0f;
              }
              *reinterpret_cast<unsigned short*>(dstp + a2_rel[jj] + 1024 * b) = p64_op_bits<ActT>(y);
            }
          }
        }
        tc_fence_before();
        fence_proxy_async();
        if (warp == 0 && lane == 0) trace_ev(a.trace, s, 3);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(d1_empty);
          mbar_arrive(&a2_full[s & 1]);
        }
      }
      if (s >= 1) {
        // ---- fin(s-1): conv2 out[t] = E[t] + L[t + 1] -> + b2 + residual terms -> fp32 stream
        const int i = s - 1;
        const int n = n_next;
        float* op = a.out + base_next;
        ewait(d2_full, (uint32_t)i & 1u);
        tc_fence_after();
        if (warp == 0 && lane == 0) trace_ev(a.trace, i, 4);
        if (!(a.dbg & 4)) {
          const bool full = 64 * tq + 64 <= n;
          const int nrel = n - col_lo;                                       // columns (relative to col_lo) inside the tile
          uint32_t v1[2][8], v2[2][8];
          ld_block(tmem_base + 256u, 64 * tq, 1, v1[0], v2[0]);
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            tmem_ld_wait();
            if (b < 7) ld_block(tmem_base + 256u, 64 * tq + 8 * (b + 1), 1, v1[(b + 1) & 1], v2[(b + 1) & 1]);
            float sum[4];
            combine(v1[b & 1], v2[b & 1], sum);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              if (full || 8 * b + jj < nrel) {
                const float o = fmaf(sum[jj], ws2, b2 + xr[4 * b + jj]) * a.out_scale;
                if (!(a.dbg & 64)) op[(8 * b + jj) * C] = o;
                if constexpr (kOact) {
                  const float sl = a.act[0].slope;
                  reinterpret_cast<unsigned short*>(a.act[0].dst)[base_next + (8 * b + jj) * C] = p64_op_bits<ActT>(o > 0.f ? o : o * sl);
                }
              }
            }
          }
        }
        tc_fence_before();
        if (warp == 0 && lane == 0) trace_ev(a.trace, i, 5);
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty);
        // residual terms of the next tile: in flight during mid(s+1) and the wait for conv2
        if (i + 1 < n_my) {
          base_next = tile_base(i + 1, n_next);
          issue_x(base_next, n_next);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kP64W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int EM, int NA1, typename ActT>
int launch_pair64_na(const ConvArgs& a, const PairArgs& p, uint32_t idesc, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    VT_CUDA_OK(cudaFuncSetAttribute(k_pair64_tc<EM, NA1, ActT>, cudaFuncAttributeMaxDynamicSharedMemorySize, P64Cfg<NA1>::kSmem));
    configured = true;
  }
  VT_CUDA_OK(launch_pdl(k_pair64_tc<EM, NA1, ActT>, dim3((unsigned)grid), dim3(kP64Warps * 32), (size_t)P64Cfg<NA1>::kSmem, st, a, p, idesc));
  VT_LAUNCHED();
  return VT_OK;
}

template <int EM, typename ActT>
int launch_pair64_em(const ConvArgs& a, const PairArgs& p, uint32_t idesc, int grid, cudaStream_t st) {
  static const bool na1 = getenv("VT_P64_NA1") && getenv("VT_P64_NA1")[0] == '1';
  return na1 ? launch_pair64_na<EM, 1, ActT>(a, p, idesc, grid, st) : launch_pair64_na<EM, 2, ActT>(a, p, idesc, grid, st);
}

template <typename ActT>
int launch_pair64_t(const ConvArgs& a, const PairArgs& p, uint32_t idesc, int grid, cudaStream_t st) {
  VT_REQUIRE(a.out && a.res1 && !a.act[1].dst && !a.act[2].dst, "pair64_tc: needs an fp32 output and the residual stream");
  const bool oact = a.act[0].dst && a.act[0].kind == ACT_LRELU && a.act_from_out;
  VT_REQUIRE(oact || !a.act[0].dst, "pair64_tc: only a leaky-ReLU output copy is supported");
  if (a.res2) {
    VT_REQUIRE(!oact && !a.out_accum && a.out_scale == 1.0f, "pair64_tc: unsupported epilogue with two residuals");
    return launch_pair64_em<EM_RES1 | EM_RES2 | EM_OUT, ActT>(a, p, idesc, grid, st);
  }
  if (oact) return launch_pair64_em<EM_RES1 | EM_OUT | EM_ACCUM | EM_OACT, ActT>(a, p, idesc, grid, st);
  if (a.out_accum || a.out_scale != 1.0f) return launch_pair64_em<EM_RES1 | EM_OUT | EM_ACCUM, ActT>(a, p, idesc, grid, st);
  return launch_pair64_em<EM_RES1 | EM_OUT, ActT>(a, p, idesc, grid, st);
}

}  // namespace tc

bool pair64_tc_supported(const ConvLayer& c1, const ConvLayer& c2) {
  return c1.w_tp && c2.w_tp && c1.cin == 64 && c1.cout == 64 && c2.cin == 64 && c2.cout == 64 && c1.k == c2.k &&
         (c1.k & 1) && c1.k <= 11 && c2.dil == 1 && c1.dil <= 5 && (c1.k - 1) * c1.dil <= tc::kP64RA1 - 256 - 6 && c1.stride == 1 &&
         c2.stride == 1 && c1.out_mul == 1 && c2.out_mul == 1;
}

// Output steps per tile: conv1 columns < 248 are valid, conv2 column t needs the conv1 columns t .. t + k - 1.
int pair64_tc_tile_rows(int k) { return 249 - k; }

// Tap-pair image of a C = 64 layer: one 16 KB slab (128 rows x 64 input channels, SWIZZLE_128B) per pair of taps
// (2p, 2p+1) - the last pair of an odd kernel has a zero L half.  Row 32q + l of a slab is output channel
// 16q + (l & 15) of tap 2p (l < 16) or 2p + 1 (l >= 16); source and destination rows agree modulo 8, so the
// swizzled 128-byte rows of the pack_conv_tc image are copied verbatim.
int pack_pair64(ConvLayer& L, std::vector<void*>& allocs) {
  if (!L.w_tc || L.cin != 64 || L.cout != 64 || L.stride != 1 || L.out_mul != 1) return VT_OK;
  const int npairs = (L.k + 1) / 2;
  std::vector<uint8_t> src((size_t)L.k * 8192), img((size_t)npairs * 16384, 0);
  VT_CUDA_OK(cudaMemcpy(src.data(), L.w_tc, src.size(), cudaMemcpyDeviceToHost));
  for (int pp = 0; pp < npairs; ++pp)
    for (int m = 0; m < 128; ++m) {
      const int q = m >> 5, l = m & 31;
      const int tap = 2 * pp + (l >= 16 ? 1 : 0), co = 16 * q + (l & 15);
      if (tap < L.k) std::memcpy(img.data() + (size_t)pp * 16384 + (size_t)m * 128, src.data() + (size_t)tap * 8192 + (size_t)co * 128, 128);
    }
  void* p = nullptr;
  VT_CUDA_OK(cudaMalloc(&p, img.size()));
  allocs.push_back(p);
  VT_CUDA_OK(cudaMemcpy(p, img.data(), img.size(), cudaMemcpyHostToDevice));
  L.w_tp = p;
  return VT_OK;
}

int launch_pair64_tc(const ConvArgs& a_in, const ConvLayer& c1, const ConvLayer& c2, const float* alpha1, const float* alpha2,
                     int act_elem, cudaStream_t st) {
  VT_REQUIRE(pair64_tc_supported(c1, c2) && (act_elem == ELEM_F16 || act_elem == ELEM_BF16), "pair64_tc: layers %s / %s cannot be fused",
             c1.name.c_str(), c2.name.c_str());
  if (a_in.n_tiles == 0) return VT_OK;
  ConvArgs a = a_in;
  static const int dbg = getenv("VT_TC_DBG") ? atoi(getenv("VT_TC_DBG")) : 0;
  a.dbg = dbg;
  a.bias = c2.bias;
  a.cout = c2.cout; a.phase_c = c2.cout; a.out_mul = 1; a.out_shift = 0; a.dup_row2 = 0;
  PairArgs p{};
  p.x_in = a.res1;
  p.alpha1 = alpha1; p.alpha2 = alpha2; p.bias1 = c1.bias; p.wscale1 = c1.wscale;
  a.wscale = c2.wscale;
  p.w1 = reinterpret_cast<const uint8_t*>(c1.w_tp);
  p.w2 = reinterpret_cast<const uint8_t*>(c2.w_tp);
  p.k = c1.k; p.dil = c1.dil;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    VT_CUDA_OK(cudaGetDevice(&dev));
    VT_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
  static const char* trace_name = getenv("VT_TC_TRACE");
  static long long* d_trace = nullptr;
  const bool tracing = trace_name && c1.name == trace_name;
  if (tracing) {
    if (!d_trace) VT_CUDA_OK(cudaMalloc(&d_trace, tc::kTraceTiles * tc::kTraceEvents * 8));
    VT_CUDA_OK(cudaMemsetAsync(d_trace, 0, tc::kTraceTiles * tc::kTraceEvents * 8, st));
    a.trace = d_trace;
  }
  const uint32_t fmt = act_elem == ELEM_F16 ? 0u : 1u;
  // D fp32, A / B fp16 or bf16, both K-major, N = 256 (time steps), M = 128 (two taps x 64 output channels)
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
  const int rc = act_elem == ELEM_F16 ? tc::launch_pair64_t<__half>(a, p, idesc, grid, st)
                                      : tc::launch_pair64_t<__nv_bfloat16>(a, p, idesc, grid, st);
  if (tracing && rc == VT_OK) {
    std::vector<long long> h(tc::kTraceTiles * tc::kTraceEvents);
    VT_CUDA_OK(cudaStreamSynchronize(st));
    VT_CUDA_OK(cudaMemcpy(h.data(), d_trace, h.size() * 8, cudaMemcpyDeviceToHost));
    long long t0 = 0;
    for (size_t q = 0; q < h.size(); ++q)
      if ((int)(q % tc::kTraceEvents) < 10 && h[q] && (!t0 || h[q] < t0)) t0 = h[q];
    fprintf(stderr, "[vt trace] pair64 %s k=%d dil=%d tiles=%d grid=%d (cycles; PROD start end | MID start end | FIN start end | "
            "C1 start issued | C2 start issued)\n", c1.name.c_str(), c1.k, c1.dil, a.n_tiles, grid);
    for (int it = 0; it < tc::kTraceTiles; ++it) {
      if (!h[it * tc::kTraceEvents + 0]) break;
      fprintf(stderr, "[vt trace] %2d", it);
      for (int e = 0; e < 10; ++e) fprintf(stderr, " %7lld", h[it * tc::kTraceEvents + e] ? h[it * tc::kTraceEvents + e] - t0 : -1);
      fprintf(stderr, "  w_wait c1=%lld c2=%lld x_wait=%lld\n", h[it * tc::kTraceEvents + 10], h[it * tc::kTraceEvents + 11], h[it * tc::kTraceEvents + 12]);
    }
  }
  return rc;
}

}  // namespace vt
