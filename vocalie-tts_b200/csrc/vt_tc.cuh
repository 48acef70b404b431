// tcgen05 / TMEM / mbarrier / bulk-copy primitives and the shared convolution epilogue of the
// tensor-core kernels (vt_conv_tc.cu: activation-resident ResBlock convs; vt_gemm_tc.cu: K-blocked
// plain layers).  sm_100a only.
#pragma once
#include "vt_hift.cuh"

#include <cstdlib>
#include <utility>

namespace vt {

// One ResBlock iteration run as a fused pair (vt_pair_tc.cu, vt_pair64_tc.cu)
struct PairArgs {
  const float* x_in;      // fp32 residual stream [rows][C]; gap rows are zero
  const float* alpha1;    // Snake before conv1
  const float* alpha2;    // Snake before conv2
  const float* bias1;     // conv1 bias
  const float* wscale1;   // conv1 inverse weight-row scales (ConvArgs::wscale holds conv2's)
  const uint8_t* w1;      // conv1 / conv2 weights: pack_conv_tc images (chunk = (tap, 64-channel block)), or the
  const uint8_t* w2;      // tap-pair images of pack_pair64
  int k, dil;
  // Mean-fused launch (nsub == 3): the LAST pairs of the three ResBlocks of a stage run as sub-iterations of one tile
  // and accumulate in the same TMEM buffer - out = (1/3) sum_r [x_r + b2_r + conv2_r(Snake(conv1_r(Snake(x_r))))] - so the
  // partial mean never travels through HBM.  Sub 0 is described by the fields above (and ConvArgs::bias / res1).
  int nsub;
  struct Sub {
    const float* x_in; const float* alpha1; const float* alpha2; const float* bias1; const float* bias2;
    const uint8_t* w1; const uint8_t* w2;
    int k, dil;
  } more[2];
};

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per WARP: every lane has finished its part (and issued its own proxy / tcgen05 fence), the warp
// converges, lane 0 arrives.  Barriers counted in warps see 32x fewer arrivals - and their waiters fewer wake-ups.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must fault the launch, not hang the GPU.  The clock is only read
// once the first probe has failed.
// try_wait is given a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of returning at once.  Without it a waiting warp spins through try_wait / clock / branch -
// measured on the C = 64 pair kernel (ncu source view): 43 % of all executed warp instructions were such spins,
// issue slots and power taken from the warps that work.
constexpr uint32_t kSuspendHintNs = 20000;
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity), "r"(kSuspendHintNs)
      : "memory");
  return done != 0;
}
// Slow path.  ncu's source view of the round-1 build showed 37-61 % of ALL executed warp instructions of the pair
// kernels in this function: the hardware suspend of try_wait ends early (30-70 wake-ups per wait), and every
// iteration also read the clock and did a 64-bit compare (11 instructions).  The bound is now an iteration count
// checked in an outer loop: an iteration is try_wait + sleep + recheck + count + branch.
#ifdef VT_AB_OLD_WAIT   // A/B builds only (tools/ab_build.sh): the round-1 loop with a clock read per probe
static __device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try(addr, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
#else
static __device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
#pragma unroll 1
  for (int outer = 0; outer < (1 << 14); ++outer) {
#pragma unroll 1
    for (int i = 0; i < 2048; ++i)
      if (mbar_try(addr, parity)) return;
  }
  __trap();   // ~3e7 failed probes (seconds): a protocol bug must fault the launch, not hang the GPU
}
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (!mbar_try(addr, parity)) mbar_wait_slow(addr, parity);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Same copy delivered to the same CTA-relative offset of every CTA in `mask` (cluster ranks); each destination's
// mbarrier at the offset of `bar` receives the complete_tx.
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Programmatic dependent launch: consecutive tensor-core kernels of a forward are launched with programmatic stream
// serialisation, every CTA signals at once that its dependents may be scheduled, and every role that reads what the
// previous kernel wrote waits for that kernel's completion - AFTER barrier initialisation, TMEM allocation and parameter
// staging, while the weight loader (static data) already fills its ring.  The persistent kernels hold an SM each, so the
// next kernel's CTA starts on an SM the moment this kernel's CTA leaves it: launch latency, prologue and the tail of the
// slowest SMs overlap (measured gap between dependent launches: ~12 us x 68 launches per forward).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE (canonical layout ((8,n),2):((16B,SBO),LBO)):
// 8 rows x 16 B core matrices; SBO = byte distance between 8-row groups, LBO = between the two
// 8-element K halves of one UMMA_K = 16 step.  Bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version = 1 (sm_100), [61,64) layout = 0.
// Operand layout switch.  1: K-major SWIZZLE_128B (rows of 64 elements = 128 B, 16-byte chunk c of row r
// stored at chunk c ^ (r & 7); 8-row atoms of 1024 B, SBO = 1024): the layout the tensor core fetches at
// full rate.  The swizzle is a function of the shared-memory address bits, so a descriptor whose start
// address is advanced by whole rows (128 B each) still addresses the right chunks: conv taps remain pure
// descriptor shifts.  0: the no-swizzle panels described above (kept for reference; ~4x slower operand fetch).
#ifndef VT_TC_SWIZZLE
#define VT_TC_SWIZZLE 1
#endif
constexpr bool kSwz = VT_TC_SWIZZLE != 0;

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
  // [0,14) addr>>4, [16,30) LBO (unused for swizzled K-major; 1), [32,46) SBO = 1024>>4, [46,48) version 1,
  // [61,64) layout type 2 = SWIZZLE_128B
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same MMA with the descriptors given as (lo, hi) words: the hi words are compile-time constants of the
// layout, the lo words advance by plain adds (address >> 4).
__device__ __forceinline__ void umma_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The arrive is delivered to the barrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <typename T> struct Pack8;
template <> struct Pack8<__half> {
  static __device__ __forceinline__ uint4 pack(const float* y) {
    __half2 a = __floats2half2_rn(y[0], y[1]), b = __floats2half2_rn(y[2], y[3]);
    __half2 c = __floats2half2_rn(y[4], y[5]), d = __floats2half2_rn(y[6], y[7]);
    return make_uint4(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b),
                      *reinterpret_cast<unsigned*>(&c), *reinterpret_cast<unsigned*>(&d));
  }
};
template <> struct Pack8<__nv_bfloat16> {
  static __device__ __forceinline__ uint4 pack(const float* y) {
    __nv_bfloat162 a = __floats2bfloat162_rn(y[0], y[1]), b = __floats2bfloat162_rn(y[2], y[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(y[4], y[5]), d = __floats2bfloat162_rn(y[6], y[7]);
    return make_uint4(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b),
                      *reinterpret_cast<unsigned*>(&c), *reinterpret_cast<unsigned*>(&d));
  }
};

template <typename T> struct Pack4;
template <> struct Pack4<__half> {
  static __device__ __forceinline__ uint2 pack(const float* y) {
    __half2 a = __floats2half2_rn(y[0], y[1]), b = __floats2half2_rn(y[2], y[3]);
    return make_uint2(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b));
  }
};
template <> struct Pack4<__nv_bfloat16> {
  static __device__ __forceinline__ uint2 pack(const float* y) {
    __nv_bfloat162 a = __floats2bfloat162_rn(y[0], y[1]), b = __floats2bfloat162_rn(y[2], y[3]);
    return make_uint2(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b));
  }
};

// Host side: launch with the programmatic-serialisation attribute (VT_PDL=0 restores plain launches for A/B timing).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool on = !(getenv("VT_PDL") && getenv("VT_PDL")[0] == '0');
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

constexpr int kStageLd = 36;   // floats per staged row (32 + 4 pad: conflict-free 16-byte accesses both ways)

// Snake x + sin^2(alpha x) / alpha on the SFU: sin.approx takes the angle in revolutions after one
// multiply, and the hardware reduces the range exactly, so the only error is the rounding of
// alpha*x/2pi: ~4e-7 * |alpha x| radians, orders below the fp16 rounding of the operand it feeds.
__device__ __forceinline__ float snake_f(float v, float alpha, float inv_alpha) {
  const float s = __sinf(v * alpha);
  return fmaf(inv_alpha, s * s, v);
}

// Two Snakes at once on the packed fp32 pipe (sm_100 FMUL2 / FFMA2: one issue slot for two lanes of a register pair).
// Same operations and roundings as two snake_f calls - bit-identical results - with 3.5 instead of 5 instructions per
// element (the two range-reduction multiplies and the two MUFU.SIN stay scalar).  The pair kernels' producer and mid roles
// are bound by the issue rate of exactly this instruction stream.
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
#ifdef VT_NO_F32X2            // A/B build switch (tools/ab_build.sh scalar -DVT_NO_F32X2): the same arithmetic on the scalar pipe
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  float a0, a1, b0, b1; f2_unpack(a, a0, a1); f2_unpack(b, b0, b1);
  return f2_pack(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  float a0, a1, b0, b1; f2_unpack(a, a0, a1); f2_unpack(b, b0, b1);
  return f2_pack(__fadd_rn(a0, b0), __fadd_rn(a1, b1));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  float a0, a1, b0, b1, c0, c1; f2_unpack(a, a0, a1); f2_unpack(b, b0, b1); f2_unpack(c, c0, c1);
  return f2_pack(__fmaf_rn(a0, b0, c0), __fmaf_rn(a1, b1, c1));
}
#else
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
#endif
// y = v + inv_alpha * sin^2(alpha v) on a packed pair (v, alpha, inv_alpha: packed pairs)
__device__ __forceinline__ unsigned long long snake_f2(unsigned long long v, unsigned long long alpha, unsigned long long inv_alpha) {
  float t0, t1;
  f2_unpack(f2_mul(v, alpha), t0, t1);
  const unsigned long long s = f2_pack(__sinf(t0), __sinf(t1));
  return f2_fma(inv_alpha, f2_mul(s, s), v);
}
__device__ __forceinline__ void snake_f2(float v0, float v1, float a0, float a1, float i0, float i1, float& y0, float& y1) {
  f2_unpack(snake_f2(f2_pack(v0, v1), f2_pack(a0, a1), f2_pack(i0, i1)), y0, y1);
}

// Debug timeline (VT_TC_TRACE): CTA 0 records clock64 at role events of its first kTraceTiles tiles.
constexpr int kTraceTiles = 48, kTraceEvents = 14;   // 0-9 role timestamps, 10/11 weight-ring waits of conv1/conv2, 12 producer x-ring wait, 13 spare
__device__ __forceinline__ void trace_ev(long long* trace, int it, int ev) {
  if (trace && blockIdx.x == 0 && it < kTraceTiles) trace[it * kTraceEvents + ev] = clock64();
}

constexpr int kProdWarps = 4;            // activation-producer warps
constexpr int kProd = kProdWarps * 32;

constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// Epilogue modes (compile-time specialisations of the shared ConvArgs epilogue)
//   v = acc + bias (+res1) (+res2); o = v*scale (+ out if ACCUM and a.out_accum); out = o
//   ACT1/ACT3: act[s] = Snake(v);  OACT: act[0] = leaky_relu(o) (operand copy for the next stage)
constexpr int EM_RES1 = 1, EM_RES2 = 2, EM_OUT = 4, EM_ACCUM = 8, EM_ACT1 = 16, EM_ACT3 = 32, EM_OACT = 64;
//   ELU: v = elu(acc + bias) first (F0 predictor trunk);  SPLIT: act[0] = fp16(v), act[1] = fp16(v - act[0])
//   (two-term operand split that keeps ~22 mantissa bits through the tensor core)
constexpr int EM_ELU = 128, EM_SPLIT = 256;

// Epilogue of one accumulator tile (MB x 128 rows x NT columns, fp32 in TMEM at `tmem_acc`), run by the
// NEPI epilogue warps.  Warp w may touch TMEM lanes 32*(w%4)..+31; the NEPI/4 warps of a lane quarter
// split the tile's 32-column blocks round-robin.  tcgen05.ld gives every thread 32 columns of ITS row;
// global memory wants a warp to touch whole rows, so each 32x32 fp32 block is transposed through a
// padded shared-memory stage and every global instruction covers 4 rows x 128 contiguous bytes.
template <int NT, int MB, int NEPI, int EM, typename ActT>
__device__ __forceinline__ void epilogue_tile(const ConvArgs& a, const ConvTile& tile, const int nt, const uint32_t tmem_acc,
                                              float* stage, const int warp, const int lane) {
  constexpr int NACT = (EM & EM_ACT3) ? 3 : ((EM & EM_ACT1) ? 1 : 0);
  constexpr int NBLK = MB * (NT / 32);
  const int quarter = warp & 3, eg = warp >> 2;
  const int sub = lane & 7;          // which float4 of a 32-column block
  const int rsub = lane >> 3;        // which of the 4 rows of an iteration
  const int ld = a.phase_c;
#pragma unroll 1
  for (int blk = eg; blk < NBLK; blk += NEPI / 4) {
    if (a.dbg & 4) break;
    const int mb = blk / (NT / 32), c0 = (blk - mb * (NT / 32)) * 32;
    const int row0 = mb * 128 + quarter * 32;                    // tile-local row of this warp's lane 0
    uint32_t v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mb * NT + c0), v);
    tmem_ld_wait();
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 8; ++g)
      *reinterpret_cast<float4*>(stage + lane * kStageLd + g * 4) =
          make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                      __uint_as_float(v[4 * g + 3]));
    __syncwarp();
    if (a.dbg & 8) continue;
    const int cg = nt * NT + c0 + sub * 4;             // GEMM column of this lane's float4
    const int phase = cg / ld, co = cg - phase * ld;   // transposed convs: column -> (output phase, channel)
    const float4 bias = *reinterpret_cast<const float4*>(a.bias + cg);
    const float4 wsc = *reinterpret_cast<const float4*>(a.wscale + cg);
    float4 al[NACT > 0 ? NACT : 1], ia[NACT > 0 ? NACT : 1];
#pragma unroll
    for (int s = 0; s < NACT; ++s) {
      al[s] = *reinterpret_cast<const float4*>(a.act[s].alpha + co);
      ia[s] = make_float4(__fdividef(1.0f, al[s].x + 1e-9f), __fdividef(1.0f, al[s].y + 1e-9f),
                          __fdividef(1.0f, al[s].z + 1e-9f), __fdividef(1.0f, al[s].w + 1e-9f));
    }
    // phase 1: issue every global read of this block (8 rows per lane) before any use, so the
    // loads overlap instead of serialising behind the per-row control flow
    const long long step0 = (long long)(tile.q0 + row0 + rsub) * a.out_mul + phase + a.out_shift;
    const long long idx0 = (tile.out_row0 + step0) * ld + co;
    const long long istride = 4LL * a.out_mul * ld;        // 4 rows per iteration
    const int nvalid = tile.n - row0 - rsub;                // rows r = 4i + rsub are valid while 4i < nvalid
    constexpr bool kLoads = (EM & (EM_RES1 | EM_RES2 | EM_ACCUM)) != 0;
    float4 pre[kLoads ? 8 : 1];
    if constexpr (kLoads) {
      const bool accum = ((EM & EM_ACCUM) != 0) && a.out_accum;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * i < nvalid && !(a.dbg & 32)) {
          const long long idx = idx0 + i * istride;
          if constexpr ((EM & EM_RES1) != 0) q = __ldg(reinterpret_cast<const float4*>(a.res1 + idx));
          if constexpr ((EM & EM_RES2) != 0) {
            const float4 q2 = __ldg(reinterpret_cast<const float4*>(a.res2 + idx));
            q.x += q2.x; q.y += q2.y; q.z += q2.z; q.w += q2.w;
          }
          if (accum) {
            // fold the previous partial mean into the residual term: (v + res)*s + prev = (v + res + prev/s)*s
            const float4 pv = *reinterpret_cast<const float4*>(a.out + idx);
            const float inv = 1.0f / a.out_scale;
            q.x = fmaf(pv.x, inv, q.x); q.y = fmaf(pv.y, inv, q.y); q.z = fmaf(pv.z, inv, q.z); q.w = fmaf(pv.w, inv, q.w);
          }
        }
        pre[i] = q;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + rsub;
      if (4 * i >= nvalid) continue;
      const long long step = step0 + 4LL * i * a.out_mul;
      const long long idx = idx0 + i * istride;
      const float4 acc = *reinterpret_cast<const float4*>(stage + r * kStageLd + sub * 4);
      float x[4] = {fmaf(acc.x, wsc.x, bias.x), fmaf(acc.y, wsc.y, bias.y), fmaf(acc.z, wsc.z, bias.z), fmaf(acc.w, wsc.w, bias.w)};
      if constexpr (kLoads) {
        x[0] += pre[i].x; x[1] += pre[i].y; x[2] += pre[i].z; x[3] += pre[i].w;
      }
      if constexpr ((EM & EM_ELU) != 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = x[e] > 0.0f ? x[e] : expm1f(x[e]);
      }
      if constexpr ((EM & EM_OUT) != 0) {
        float4 o = make_float4(x[0] * a.out_scale, x[1] * a.out_scale, x[2] * a.out_scale, x[3] * a.out_scale);
        if (!(a.dbg & 64)) *reinterpret_cast<float4*>(a.out + idx) = o;
        if (a.dup_row2 && step == 2)                      // reflection pad (1, 0): padded[0] = unpadded[1]
          *reinterpret_cast<float4*>(a.out + tile.out_row0 * ld + co) = o;
        if constexpr ((EM & EM_OACT) != 0) {
          const float sl = a.act[0].slope;
          float y[4] = {o.x > 0.f ? o.x : o.x * sl, o.y > 0.f ? o.y : o.y * sl, o.z > 0.f ? o.z : o.z * sl,
                        o.w > 0.f ? o.w : o.w * sl};
          *reinterpret_cast<uint2*>(reinterpret_cast<ActT*>(a.act[0].dst) + idx) = Pack4<ActT>::pack(y);
        }
      }
      if constexpr ((EM & EM_SPLIT) != 0) {
        // hi = fp16(x), lo = fp16(x - hi): both written in fp16 whatever the operand type of the vocoder
        const uint2 hi = Pack4<__half>::pack(x);
        const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&hi.x));
        const float2 h23 = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
        const float lo[4] = {x[0] - h01.x, x[1] - h01.y, x[2] - h23.x, x[3] - h23.y};
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.act[0].dst) + idx) = hi;
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.act[1].dst) + idx) = Pack4<__half>::pack(lo);
      }
#pragma unroll
      for (int s = 0; s < NACT; ++s) {
        float y[4];
        y[0] = snake_f(x[0], al[s].x, ia[s].x); y[1] = snake_f(x[1], al[s].y, ia[s].y);
        y[2] = snake_f(x[2], al[s].z, ia[s].z); y[3] = snake_f(x[3], al[s].w, ia[s].w);
        *reinterpret_cast<uint2*>(reinterpret_cast<ActT*>(a.act[s].dst) + idx) = Pack4<ActT>::pack(y);
      }
    }
  }
}

}  // namespace tc
}  // namespace vt
