// Post-processing kernels: silence trim, zero-cross snap, fades, peak normalise, gap-padded
// concatenation, PCM_16.  Bandwidth-bound byte/float work: 128-bit coalesced HBM access,
// warp-shuffle reductions, one atomic per block.  Bit-exact against the reference's numpy
// (backend/shared/tts_pipeline.py:114-274, backend/shared/audio_edit.py:16-79): all float
// arithmetic uses explicit round-to-nearest intrinsics so nothing is contracted into FMAs.
//
// Passes (algorithmic HBM bytes per input sample, fp32 in / fp32 out):
//   k_scan  : read 4 B  -> first/last active index AND per-tile max|x|   (one fused read pass)
//   k_fix   : per segment, O(radius) samples: min-silence rule, snap, fade lengths
//   k_peak  : only the tiles cut by the trim range or a fade are re-read (O(fade) samples);
//             interior tiles reuse the per-tile maxima from k_scan
//   k_plan  : per segment scale (float64 divide) + exclusive scan of output offsets
//   k_write : read 4 B + write 4 B (2 B for PCM_16)
// => 12 B/sample (10 B with PCM_16 output), as stated in DESIGN.md.
#include "vt_common.cuh"

#include <climits>

namespace vt {

constexpr int kThreads = 256;
constexpr int kScanTile = 1024;    // samples per warp iteration of the scan pass (8 x float4 per lane)
constexpr int kWriteTile = 4096;   // output samples per block iteration (16 per thread)

struct __align__(16) SegPlan {
  long long first;       // first active sample (segment-relative); LLONG_MAX if none
  long long last;        // last active sample; -1 if none
  long long start, end;  // final range after min-silence / snap / fallback
  long long dst;         // output offset (samples)
  long long out_len;     // end - start
  int fi, fo;            // clipped fade-in / fade-out lengths
  unsigned peak_bits;    // float bits of max|x| over the processed segment
  float scale;
  int apply_scale;
  int gap_after;
  long long tile_base;   // index of this segment's first entry in the tile-max array
};

struct PostHeader {
  unsigned global_peak_bits;
  int pad;
  long long total_out;
};

// ---- reference ramps: np.linspace(0,1,F) / np.linspace(1,0,F) cast to float32 -------------
__device__ __forceinline__ float ramp_in(int F, long long j) {
  if (F == 1) return 0.0f;
  if (j == F - 1) return 1.0f;
  const double step = __ddiv_rn(1.0, (double)(F - 1));
  return __double2float_rn(__dmul_rn((double)j, step));
}
__device__ __forceinline__ float ramp_out(int F, long long j) {
  if (F == 1) return 1.0f;
  if (j == F - 1) return 0.0f;
  const double step = __ddiv_rn(-1.0, (double)(F - 1));
  return __double2float_rn(__dadd_rn(__dmul_rn((double)j, step), 1.0));
}
// j is relative to the trimmed segment of length len.
__device__ __forceinline__ float apply_fades(float x, long long j, long long len, int fi, int fo,
                                             int out_first) {
  const bool zin = j < fi;
  const bool zout = j >= len - fo;
  if (!(zin | zout)) return x;
  if (out_first) {
    if (zout) x = __fmul_rn(x, ramp_out(fo, j - (len - fo)));
    if (zin) x = __fmul_rn(x, ramp_in(fi, j));
  } else {
    if (zin) x = __fmul_rn(x, ramp_in(fi, j));
    if (zout) x = __fmul_rn(x, ramp_out(fo, j - (len - fo)));
  }
  return x;
}

__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0f;
    t = warp_max(t);
    if (threadIdx.x == 0) sh[0] = t;
  }
  __syncthreads();
  float r = sh[0];
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------------------------ init
__global__ void k_post_init(SegPlan* plan, PostHeader* hdr, const int64_t* seg_off, int n_seg,
                            const int64_t* range_override) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    hdr->global_peak_bits = 0u;
    hdr->total_out = 0;
  }
  if (i >= n_seg) return;
  SegPlan& p = plan[i];
  p.first = LLONG_MAX;
  p.last = -1;
  p.peak_bits = 0u;
  p.scale = 1.0f;
  p.apply_scale = 0;
  if (range_override) {
    p.start = range_override[2 * i];
    p.end = range_override[2 * i + 1];
  }
}

// Block-wide exclusive scan of one long long per thread (blockDim.x a multiple of 32, <= 1024); `carry` is a
// shared running total carried across calls.  Returns this thread's exclusive prefix including the carry.
__device__ __forceinline__ long long block_excl_scan(long long v, long long* sh_warp, long long* sh_carry) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sh_warp[wid] = incl;
  __syncthreads();
  long long woff = 0;
  for (int w = 0; w < wid; ++w) woff += sh_warp[w];
  const long long carry = *sh_carry;
  __syncthreads();
  if (threadIdx.x == blockDim.x - 1) *sh_carry = carry + woff + incl;
  __syncthreads();
  return carry + woff + incl - v;
}

// Exclusive scan of per-segment tile counts (one block).
__global__ void k_tile_bases(SegPlan* plan, const int64_t* seg_off, int n_seg) {
  __shared__ long long sh_warp[32];
  __shared__ long long sh_carry;
  if (threadIdx.x == 0) sh_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_seg; base += blockDim.x) {
    const int i = base + threadIdx.x;
    long long v = 0;
    if (i < n_seg) {
      const long long a = seg_off[i], b = seg_off[i + 1];
      const long long A = a & ~3LL;
      v = (b - A + kScanTile - 1) / kScanTile;
    }
    const long long ex = block_excl_scan(v, sh_warp, &sh_carry);
    if (i < n_seg) plan[i].tile_base = ex;
  }
}

// ------------------------------------------------------------------------------------ scan
// One read pass: first/last index with |x| > thr, and max|x| of every tile.  A tile is the
// unit of one warp iteration (kScanTile samples = 8 independent 128-bit loads per lane, no
// block barrier in the loop); tiles are aligned to 16 B in *absolute* buffer coordinates so
// every load is a full float4.
__global__ void __launch_bounds__(kThreads)
k_scan(const float* __restrict__ audio, const int64_t* __restrict__ seg_off, int n_seg,
       float thr, int do_range, SegPlan* plan, float* __restrict__ tile_max) {
  __shared__ long long sh_lo[kThreads / 32], sh_hi[kThreads / 32];
  const int seg = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long a = seg_off[seg], b = seg_off[seg + 1];
  const long long n_total = seg_off[n_seg];
  if (b <= a) return;
  const long long A = a & ~3LL;
  const long long n_tiles = (b - A + kScanTile - 1) / kScanTile;
  const long long tbase = plan[seg].tile_base;
  long long lo = LLONG_MAX, hi = -1;
  for (long long tile = (long long)blockIdx.x * (kThreads / 32) + warp; tile < n_tiles;
       tile += (long long)gridDim.x * (kThreads / 32)) {
    const long long t0 = A + tile * kScanTile;
    float4 q[kScanTile / 128];
    const bool fast = (t0 + kScanTile <= n_total);
    if (fast) {
#pragma unroll
      for (int u = 0; u < kScanTile / 128; ++u) q[u] = ldg_stream4(audio + t0 + (u * 32 + lane) * 4);
    } else {
#pragma unroll
      for (int u = 0; u < kScanTile / 128; ++u) {
        const long long p = t0 + (u * 32 + lane) * 4;
        q[u].x = (p + 0 < n_total) ? audio[p + 0] : 0.0f;
        q[u].y = (p + 1 < n_total) ? audio[p + 1] : 0.0f;
        q[u].z = (p + 2 < n_total) ? audio[p + 2] : 0.0f;
        q[u].w = (p + 3 < n_total) ? audio[p + 3] : 0.0f;
      }
    }
    float m = 0.0f;
    const bool inside = (t0 >= a) && (t0 + kScanTile <= b);
    // bit (4u + e) of `act`: element e of this lane's u-th float4 is inside the segment and above the threshold.
    // Two or three instructions per sample on the common path; the 64-bit index arithmetic runs once per tile.
    unsigned act = 0u;
    if (inside) {
#pragma unroll
      for (int u = 0; u < kScanTile / 128; ++u) {
        const float v[4] = {fabsf(q[u].x), fabsf(q[u].y), fabsf(q[u].z), fabsf(q[u].w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          m = fmaxf(m, v[e]);
          act |= (v[e] > thr ? 1u : 0u) << (4 * u + e);
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < kScanTile / 128; ++u) {
        const long long p = t0 + (u * 32 + lane) * 4;
        const float v[4] = {fabsf(q[u].x), fabsf(q[u].y), fabsf(q[u].z), fabsf(q[u].w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (p + e >= a && p + e < b) {
            m = fmaxf(m, v[e]);
            act |= (v[e] > thr ? 1u : 0u) << (4 * u + e);
          }
        }
      }
    }
    if (act) {
      const int bl = __ffs(act) - 1, bh = 31 - __clz(act);
      const long long rl = t0 + ((bl >> 2) * 32 + lane) * 4 + (bl & 3) - a;
      const long long rh = t0 + ((bh >> 2) * 32 + lane) * 4 + (bh & 3) - a;
      lo = rl < lo ? rl : lo;
      hi = rh > hi ? rh : hi;
    }
    m = warp_max(m);
    if (lane == 0) tile_max[tbase + tile] = m;
  }
  if (!do_range) return;
  lo = warp_min_ll(lo);
  hi = warp_max_ll(hi);
  if (lane == 0) {
    sh_lo[warp] = lo;
    sh_hi[warp] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) {
      lo = sh_lo[w] < lo ? sh_lo[w] : lo;
      hi = sh_hi[w] > hi ? sh_hi[w] : hi;
    }
    if (lo != LLONG_MAX) atomicMin(&plan[seg].first, lo);
    if (hi >= 0) atomicMax(&plan[seg].last, hi);
  }
}

// ------------------------------------------------------------------------------------ snap
// _snap_zero_crossing for one index; all threads of the block participate and get the result.
__device__ long long snap_block(const float* __restrict__ x, long long n, long long idx, int radius,
                                unsigned long long* sh_key) {
  if (n == 0) return idx;
  idx = idx > n - 1 ? n - 1 : idx;
  idx = idx < 0 ? 0 : idx;
  const long long lo = (idx - radius) > 1 ? (idx - radius) : 1;
  const long long hi = (idx + radius) < (n - 1) ? (idx + radius) : (n - 1);
  unsigned long long key = ~0ULL;
  for (long long i = lo + threadIdx.x; i <= hi; i += blockDim.x) {
    const float prev = x[i - 1], cur = x[i];
    const bool cross = (prev == 0.0f) || (cur == 0.0f) || (prev < 0.0f && 0.0f <= cur) ||
                       (prev > 0.0f && 0.0f >= cur);
    if (cross) {
      const unsigned long long d = (unsigned long long)(i > idx ? i - idx : idx - i);
      const unsigned long long k = (d << 32) | (unsigned long long)(i - lo);
      key = k < key ? k : key;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o);
    key = t < key ? t : key;
  }
  __syncthreads();  // protect sh_key reuse between consecutive calls
  if ((threadIdx.x & 31) == 0) sh_key[threadIdx.x >> 5] = key;
  __syncthreads();
  key = sh_key[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) key = sh_key[w] < key ? sh_key[w] : key;
  if (key == ~0ULL) return idx;
  return lo + (long long)(key & 0xffffffffULL);
}

__global__ void __launch_bounds__(kThreads)
k_snap_only(const float* __restrict__ audio, const int64_t* __restrict__ seg_off,
            const int64_t* __restrict__ idx_in, int radius, int64_t* __restrict__ idx_out) {
  __shared__ unsigned long long sh_key[kThreads / 32];
  const int seg = blockIdx.x;
  const long long a = seg_off[seg], n = seg_off[seg + 1] - a;
  const long long r = snap_block(audio + a, n, idx_in[seg], radius, sh_key);
  if (threadIdx.x == 0) idx_out[seg] = r;
}

__global__ void k_range_out(const SegPlan* plan, const int64_t* seg_off, int n_seg, int min_sil,
                            int64_t* ranges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seg) return;
  const long long n = seg_off[i + 1] - seg_off[i];
  long long s = 0, e = n;
  if (n > 0 && plan[i].last >= 0) {
    s = plan[i].first;
    e = plan[i].last + 1;
    if (s < min_sil) s = 0;
    if (n - e < min_sil) e = n;
  }
  ranges[2 * i] = s;
  ranges[2 * i + 1] = e;
}

// Raw per-segment statistics for callers that merge ranges across ranks: first / last index with |x| > thr
// (-1 / -1 when the segment has none) and max|x| over the whole segment, from k_scan's plan and tile maxima.
__global__ void __launch_bounds__(kThreads)
k_seg_stats(const SegPlan* __restrict__ plan, const int64_t* __restrict__ seg_off, const float* __restrict__ tile_max,
            int64_t* __restrict__ first_last, float* __restrict__ peak) {
  __shared__ float sh_f[kThreads / 32];
  const int seg = blockIdx.x;
  const long long a = seg_off[seg], b = seg_off[seg + 1];
  const long long A = a & ~3LL;
  const long long n_tiles = b > a ? (b - A + kScanTile - 1) / kScanTile : 0;
  const long long tbase = plan[seg].tile_base;
  float m = 0.0f;
  for (long long t = threadIdx.x; t < n_tiles; t += kThreads) m = fmaxf(m, tile_max[tbase + t]);
  m = block_max(m, sh_f);
  if (threadIdx.x == 0) {
    const bool any = plan[seg].last >= 0;
    first_last[2 * seg] = any ? plan[seg].first : -1;
    first_last[2 * seg + 1] = any ? plan[seg].last : -1;
    peak[seg] = m;
  }
}

// ------------------------------------------------------------------------------------ fix
// Per segment: min-silence rule, snap, fallback, fade lengths (one block per segment).
__global__ void __launch_bounds__(kThreads)
k_fix(const float* __restrict__ audio, const int64_t* __restrict__ seg_off, int n_seg, SegPlan* plan,
      int trim, int min_sil, int snap_radius, int fade_in, int fade_out, int stitch, int has_override,
      int head, int tail) {
  __shared__ unsigned long long sh_key[kThreads / 32];
  const int seg = blockIdx.x;
  const long long a = seg_off[seg], n = seg_off[seg + 1] - a;
  SegPlan& p = plan[seg];
  long long s = 0, e = n;
  if (has_override) {
    s = p.start;
    e = p.end;
  } else if (trim && n > 0) {
    if (p.last >= 0) {
      s = p.first;
      e = p.last + 1;
      if (s < min_sil) s = 0;
      if (n - e < min_sil) e = n;
    }
    if (snap_radius >= 0) {
      s = snap_block(audio + a, n, s, snap_radius, sh_key);
      const long long e_in = (e - 1) > s ? (e - 1) : s;
      e = snap_block(audio + a, n, e_in, snap_radius, sh_key) + 1;
    }
    if (e <= s) {
      s = 0;
      e = n;
    }
  }
  if (threadIdx.x == 0) {
    const long long len = e - s;
    int fi = fade_in < 0 ? 0 : fade_in, fo = fade_out < 0 ? 0 : fade_out;
    if ((long long)fi > len) fi = (int)len;
    if ((long long)fo > len) fo = (int)len;
    if (stitch) {
      if (seg == 0 && head) fi = 0;
      if (seg == n_seg - 1 && tail) fo = 0;
    }
    p.start = s;
    p.end = e;
    p.out_len = len;
    p.fi = fi;
    p.fo = fo;
  }
}

// ------------------------------------------------------------------------------------ peak
// max|x| of the processed (trimmed + faded) segment.  Tiles fully inside the un-faded interior
// reuse k_scan's per-tile maxima; only tiles cut by the trim range or a fade are re-read.
__global__ void __launch_bounds__(kThreads)
k_peak(const float* __restrict__ audio, const int64_t* __restrict__ seg_off, int n_seg, SegPlan* plan,
       const float* __restrict__ tile_max, int out_first, PostHeader* hdr) {
  __shared__ float sh_f[kThreads / 32];
  const int seg = blockIdx.y;
  const long long a = seg_off[seg], b = seg_off[seg + 1];
  if (b <= a) return;
  const SegPlan p = plan[seg];
  const long long len = p.out_len;
  if (len <= 0) return;
  const long long A = a & ~3LL;
  const long long n_tiles = (b - A + kScanTile - 1) / kScanTile;
  // absolute interior range [ia, ib): inside the trim range and outside both fades
  const long long ra = a + p.start, rb = a + p.end;
  const long long ia = ra + p.fi, ib = rb - p.fo;
  float m = 0.0f;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < n_tiles;
       t += (long long)gridDim.x * kThreads) {
    const long long t0 = A + t * kScanTile, t1 = t0 + kScanTile;
    // untouched interior tile (inside the trim range, outside both fades, inside [a, b)): reuse k_scan's maximum
    if (t0 >= ia && t1 <= ib) m = fmaxf(m, tile_max[p.tile_base + t]);
  }
  // The samples of every other tile that overlaps [ra, rb) - the two boundary zones cut by the trim range or a
  // fade, a few tiles at most - are re-read by the whole block 0 of the segment with coalesced strided loads
  // (one thread walking a 1024-sample tile alone made this kernel as slow as the full-bandwidth passes).
  if (blockIdx.x == 0) {
    // first interior tile boundary at or after ia, last interior tile boundary at or before ib
    long long hi_a = A + ((ia - A + kScanTile - 1) / kScanTile) * kScanTile;   // head zone [ra, hi_a)
    long long lo_b = A + ((ib - A) / kScanTile) * kScanTile;                   // tail zone [lo_b, rb)
    if (hi_a > rb) hi_a = rb;
    if (lo_b < hi_a) lo_b = hi_a;                                              // zones meet: one pass over [ra, rb)
    if (ia >= ib) { hi_a = rb; lo_b = rb; }                                    // fades overlap: no interior at all
    for (long long idx = ra + threadIdx.x; idx < hi_a; idx += kThreads)
      m = fmaxf(m, fabsf(apply_fades(audio[idx], idx - ra, len, p.fi, p.fo, out_first)));
    for (long long idx = lo_b + threadIdx.x; idx < rb; idx += kThreads)
      m = fmaxf(m, fabsf(apply_fades(audio[idx], idx - ra, len, p.fi, p.fo, out_first)));
  }
  const float bm = block_max(m, sh_f);
  if (threadIdx.x == 0 && bm > 0.0f) {
    atomicMax(&plan[seg].peak_bits, __float_as_uint(bm));
    atomicMax(&hdr->global_peak_bits, __float_as_uint(bm));
  }
}

// ------------------------------------------------------------------------------------ plan
__global__ void k_plan(SegPlan* plan, PostHeader* hdr, const int64_t* seg_off, int n_seg, int normalize,
                       double target_peak, const float* peak_override, int concat, int gap_frames, int tail,
                       double* results) {
  for (int i = threadIdx.x; i < n_seg; i += blockDim.x) {
    SegPlan& p = plan[i];
    float peak = __uint_as_float(p.peak_bits);
    if (normalize == 2) peak = __uint_as_float(hdr->global_peak_bits);
    if (peak_override) peak = *peak_override;
    double scale = 1.0;
    int apply = 0;
    if (normalize && peak > 0.0f && target_peak > 0.0) {
      scale = __ddiv_rn(target_peak, (double)peak);
      apply = 1;
    }
    p.scale = __double2float_rn(scale);
    p.apply_scale = apply;
    p.gap_after = (concat && (i < n_seg - 1 || !tail)) ? gap_frames : 0;
    if (results) {
      double* r = results + (size_t)i * VT_POST_RESULT_STRIDE;
      r[0] = (double)p.start;
      r[1] = (double)p.end;
      r[2] = (double)__uint_as_float(p.peak_bits);
      r[3] = scale;
      r[5] = (double)p.out_len;
      r[6] = (double)peak;
      r[7] = 0.0;
    }
  }
  __syncthreads();
  // exclusive scan of (out_len + gap_after) over the segments, one chunk of blockDim.x segments at a time
  __shared__ long long sh_warp[32];
  __shared__ long long sh_carry;
  if (threadIdx.x == 0) sh_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_seg; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const long long v = i < n_seg ? plan[i].out_len + plan[i].gap_after : 0;
    const long long ex = block_excl_scan(v, sh_warp, &sh_carry);
    if (i < n_seg) {
      const long long dst = concat ? ex : (long long)seg_off[i];
      plan[i].dst = dst;
      if (results) results[(size_t)i * VT_POST_RESULT_STRIDE + 4] = (double)dst;
    }
  }
  if (threadIdx.x == 0) hdr->total_out = concat ? sh_carry : (long long)seg_off[n_seg];
}

__global__ void k_total_out(const PostHeader* hdr, int64_t* total_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total_out = hdr->total_out;
}

// ------------------------------------------------------------------------------------ write
__device__ __forceinline__ short to_pcm16(float x) {
  // libsndfile default float -> PCM_16: lrintf(x * 0x7FFF), low 16 bits kept
  return (short)__float2int_rn(__fmul_rn(x, 32767.0f));
}

// One block iteration produces kWriteTile output samples whose absolute output index is
// 16 B-aligned; the (arbitrarily misaligned) input span is staged through shared memory
// with aligned 128-bit loads so that both sides of HBM see full-width coalesced accesses.
template <bool PCM16>
__global__ void __launch_bounds__(kThreads)
k_write(const float* __restrict__ audio, const int64_t* __restrict__ seg_off, int n_seg,
        const SegPlan* __restrict__ plan, int out_first, int clip, void* __restrict__ out_v,
        long long out_capacity) {
  __shared__ __align__(16) float sh_in[kWriteTile + 8];
  const int seg = blockIdx.y;
  const SegPlan p = plan[seg];
  const long long n_total = seg_off[n_seg];
  const long long span = p.out_len + p.gap_after;
  if (span <= 0) return;
  const long long d_begin = p.dst, d_end = p.dst + span;
  const long long D0 = d_begin & ~3LL;
  const long long n_tiles = (d_end - D0 + kWriteTile - 1) / kWriteTile;
  const long long src0 = seg_off[seg] + p.start;  // absolute input index of output j = 0
  float* out_f = reinterpret_cast<float*>(out_v);
  short* out_s = reinterpret_cast<short*>(out_v);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long d0 = D0 + tile * kWriteTile;          // aligned absolute output index
    const long long j0 = d0 - d_begin;                    // may be negative on the first tile
    // input span for j in [j0, j0 + kWriteTile): absolute [p0, p0 + kWriteTile)
    const long long p0 = src0 + j0;
    const long long P = p0 >= 0 ? (p0 & ~3LL) : -(((-p0) + 3) & ~3LL);
    const int shift = (int)(p0 - P);
    __syncthreads();
    for (int v = threadIdx.x; v < kWriteTile / 4 + 2; v += kThreads) {
      const long long q = P + 4LL * v;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q >= 0 && q + 3 < n_total) {
        r = ldg_stream4(audio + q);
      } else {
        if (q + 0 >= 0 && q + 0 < n_total) r.x = audio[q + 0];
        if (q + 1 >= 0 && q + 1 < n_total) r.y = audio[q + 1];
        if (q + 2 >= 0 && q + 2 < n_total) r.z = audio[q + 2];
        if (q + 3 >= 0 && q + 3 < n_total) r.w = audio[q + 3];
      }
      *reinterpret_cast<float4*>(&sh_in[4 * v]) = r;
    }
    __syncthreads();
    // interior tile (no fade, no gap, no edge of the segment or of the output buffer in it - all but a handful
    // of tiles per segment): gain and clip only, no per-sample index arithmetic
    if (j0 >= p.fi && j0 + kWriteTile <= p.out_len - p.fo && d0 + kWriteTile <= out_capacity) {
#pragma unroll
      for (int u = 0; u < kWriteTile / (4 * kThreads); ++u) {
        const int l = (u * kThreads + threadIdx.x) * 4;
        float y[4] = {sh_in[shift + l], sh_in[shift + l + 1], sh_in[shift + l + 2], sh_in[shift + l + 3]};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (p.apply_scale) y[e] = __fmul_rn(y[e], p.scale);
          if (clip) y[e] = fminf(fmaxf(y[e], -1.0f), 1.0f);
        }
        if (PCM16) *reinterpret_cast<short4*>(out_s + d0 + l) = make_short4(to_pcm16(y[0]), to_pcm16(y[1]), to_pcm16(y[2]), to_pcm16(y[3]));
        else *reinterpret_cast<float4*>(out_f + d0 + l) = make_float4(y[0], y[1], y[2], y[3]);
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < kWriteTile / (4 * kThreads); ++u) {
      const int l = (u * kThreads + threadIdx.x) * 4;  // local output index (multiple of 4)
      const long long d = d0 + l;
      if (d >= d_end) continue;
      float y[4];
      bool valid[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long j = j0 + l + e;
        valid[e] = (j >= 0) && (j < span) && (d + e < out_capacity);
        float x = 0.0f;
        if (j >= 0 && j < p.out_len) {
          x = apply_fades(sh_in[shift + l + e], j, p.out_len, p.fi, p.fo, out_first);
          if (p.apply_scale) x = __fmul_rn(x, p.scale);
          if (clip) x = fminf(fmaxf(x, -1.0f), 1.0f);
        }
        y[e] = x;
      }
      if (valid[0] && valid[1] && valid[2] && valid[3]) {
        if (PCM16) {
          short4 s4 = make_short4(to_pcm16(y[0]), to_pcm16(y[1]), to_pcm16(y[2]), to_pcm16(y[3]));
          *reinterpret_cast<short4*>(out_s + d) = s4;
        } else {
          *reinterpret_cast<float4*>(out_f + d) = make_float4(y[0], y[1], y[2], y[3]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!valid[e]) continue;
          if (PCM16) out_s[d + e] = to_pcm16(y[e]);
          else out_f[d + e] = y[e];
        }
      }
    }
  }
}

__global__ void k_pcm16_encode(const float* __restrict__ in, short* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = i; k < n; k += stride) out[k] = to_pcm16(in[k]);
}
__global__ void k_pcm16_decode(const short* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = i; k < n; k += stride) out[k] = __fdiv_rn((float)in[k], 32768.0f);
}

// Canonical 44-byte RIFF/WAVE header of a mono PCM_16 file (what Python's wave module and libsndfile write for
// sf.write(path, x, sr) with the default subtype: tts_backends/chatterbox_runner.py:152, tts_pipeline.py:409,
// audio_edit.py:70), written on the device so that a trimmed job - whose length is data dependent and lives in device
// memory - leaves the GPU as a finished file in ONE copy.
__global__ void k_wav_header(unsigned char* dst, int sr, const long long* n_dev, long long n_host) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned long long n = (unsigned long long)(n_dev ? *n_dev : n_host);
  const unsigned data = (unsigned)(n * 2ULL), riff = 36u + data, rate = (unsigned)sr, brate = rate * 2u;
  auto u32 = [&](int o, unsigned v) { dst[o] = v & 255u; dst[o + 1] = (v >> 8) & 255u; dst[o + 2] = (v >> 16) & 255u; dst[o + 3] = v >> 24; };
  auto tag = [&](int o, const char* t) { for (int i = 0; i < 4; ++i) dst[o + i] = (unsigned char)t[i]; };
  tag(0, "RIFF"); u32(4, riff); tag(8, "WAVE"); tag(12, "fmt "); u32(16, 16u);
  dst[20] = 1; dst[21] = 0;            // PCM
  dst[22] = 1; dst[23] = 0;            // mono
  u32(24, rate); u32(28, brate);
  dst[32] = 2; dst[33] = 0;            // block align
  dst[34] = 16; dst[35] = 0;           // bits per sample
  tag(36, "data"); u32(40, data);
}

// ------------------------------------------------------------------------------------ host
struct PostWs {
  PostHeader* hdr;
  SegPlan* plan;
  float* tile_max;
  int64_t tile_capacity;
};

static int64_t post_ws_bytes(int n_seg, int64_t n_samples) {
  const int64_t tiles = n_samples / kScanTile + 2 * (int64_t)n_seg + 8;
  return align_up(sizeof(PostHeader), 256) + align_up((int64_t)sizeof(SegPlan) * (n_seg + 1), 256) +
         align_up(tiles * (int64_t)sizeof(float), 256);
}

static PostWs carve(void* ws, int n_seg, int64_t n_samples) {
  PostWs w;
  char* p = reinterpret_cast<char*>(ws);
  w.hdr = reinterpret_cast<PostHeader*>(p);
  p += align_up(sizeof(PostHeader), 256);
  w.plan = reinterpret_cast<SegPlan*>(p);
  p += align_up((int64_t)sizeof(SegPlan) * (n_seg + 1), 256);
  w.tile_max = reinterpret_cast<float*>(p);
  w.tile_capacity = n_samples / kScanTile + 2 * (int64_t)n_seg + 8;
  return w;
}

static int grid_x_for(int n_seg, int64_t max_len, int tile) {
  // enough blocks to fill 148 SMs x 8 resident CTAs, capped by the work that exists
  int64_t want = (max_len + tile - 1) / tile + 1;
  int64_t cap = (148LL * 8 + n_seg - 1) / n_seg;
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace vt

using namespace vt;

// ---- RMS helper: sqrt(mean(float64(x)^2)) per segment, fixed reduction order ---------------------------
// block (p, seg) sums the 128-bit-aligned tiles p, p + P, p + 2P, ... of the segment in float64; thread
// partials are combined by a fixed shuffle tree, block partials by k_rms_final in index order.
namespace vt {
__global__ void __launch_bounds__(kThreads)
k_rms_partial(const float* __restrict__ audio, const int64_t* __restrict__ seg_off, double* __restrict__ partial) {
  __shared__ double sh[kThreads / 32];
  const int seg = blockIdx.y;
  const long long a = seg_off[seg], b = seg_off[seg + 1];
  double acc = 0.0;
  const long long A = a & ~3LL;                               // float4-aligned start
  const long long n4 = (b - A + 3) / 4;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
    const long long idx = A + i * 4;
    if (idx >= a && idx + 4 <= b) {
      const float4 v = ldg_stream4(audio + idx);
      acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    } else {
      for (int e = 0; e < 4; ++e)
        if (idx + e >= a && idx + e < b) { const double x = audio[idx + e]; acc += x * x; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
    partial[(long long)seg * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void k_rms_final(const int64_t* __restrict__ seg_off, int n_seg, const double* __restrict__ partial,
                            double* __restrict__ out) {
  const int seg = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (seg >= n_seg || (threadIdx.x & 31) != 0) return;
  double t = 0.0;
  for (int p = 0; p < VT_RMS_PARTIALS; ++p) t += partial[(long long)seg * VT_RMS_PARTIALS + p];
  const long long n = seg_off[seg + 1] - seg_off[seg];
  out[seg] = n > 0 ? sqrt(t / (double)n) : 0.0;
}
}  // namespace vt

extern "C" {

int64_t vt_post_workspace_bytes(int n_seg, int64_t n_samples) {
  if (n_seg < 0 || n_samples < 0) return VT_ERR_INVALID;
  return post_ws_bytes(n_seg, n_samples);
}

static int check_audio(const float* audio, const int64_t* seg_off, int n_seg) {
  VT_REQUIRE(n_seg >= 0 && n_seg <= 65535, "n_seg must be in [0, 65535]");
  VT_REQUIRE(seg_off != nullptr, "seg_off is NULL");
  VT_REQUIRE(audio != nullptr || n_seg == 0, "audio is NULL");
  VT_REQUIRE((reinterpret_cast<uintptr_t>(audio) & 15) == 0, "audio must be 16-byte aligned");
  return VT_OK;
}

int vt_post_analyze(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples,
                    int64_t max_seg_len, const vt_post_params* prm, const int64_t* range_override,
                    void* workspace, int64_t workspace_bytes, void* stream_v) {
  int rc = check_audio(audio, seg_off, n_seg);
  if (rc) return rc;
  VT_REQUIRE(prm != nullptr, "params is NULL");
  VT_REQUIRE(workspace_bytes >= post_ws_bytes(n_seg, n_samples), "post workspace too small");
  if (n_seg == 0) return VT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  PostWs w = carve(workspace, n_seg, n_samples);
  const int out_first = prm->stitch ? 1 : 0;
  k_post_init<<<(n_seg + 255) / 256, 256, 0, st>>>(w.plan, w.hdr, seg_off, n_seg, range_override);
  VT_LAUNCHED();
  k_tile_bases<<<1, 256, 0, st>>>(w.plan, seg_off, n_seg);
  VT_LAUNCHED();
  const bool need_scan = (prm->trim && !range_override) || prm->normalize;
  if (need_scan) {
    dim3 g(grid_x_for(n_seg, max_seg_len, kScanTile * (kThreads / 32)), n_seg);
    k_scan<<<g, kThreads, 0, st>>>(audio, seg_off, n_seg, prm->silence_threshold,
                                   (prm->trim && !range_override) ? 1 : 0, w.plan, w.tile_max);
    VT_LAUNCHED();
  }
  k_fix<<<n_seg, kThreads, 0, st>>>(audio, seg_off, n_seg, w.plan, prm->trim, prm->min_silence_frames,
                                    prm->snap_radius, prm->fade_in_frames, prm->fade_out_frames, prm->stitch,
                                    range_override ? 1 : 0, prm->stitch_head, prm->stitch_tail);
  VT_LAUNCHED();
  if (prm->normalize) {
    dim3 g(grid_x_for(n_seg, max_seg_len, kScanTile * kThreads), n_seg);
    k_peak<<<g, kThreads, 0, st>>>(audio, seg_off, n_seg, w.plan, w.tile_max, out_first, w.hdr);
    VT_LAUNCHED();
  }
  return VT_OK;
}

int vt_post_write(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples,
                  int64_t max_seg_len, const vt_post_params* prm, const float* peak_override,
                  void* out, int64_t out_capacity, double* results, int64_t* total_out,
                  void* workspace, int64_t workspace_bytes, void* stream_v) {
  int rc = check_audio(audio, seg_off, n_seg);
  if (rc) return rc;
  VT_REQUIRE(prm != nullptr, "params is NULL");
  VT_REQUIRE(out != nullptr || n_seg == 0, "out is NULL");
  VT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
  VT_REQUIRE(workspace_bytes >= post_ws_bytes(n_seg, n_samples), "post workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  if (n_seg == 0) {
    if (total_out) VT_CUDA_OK(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
    return VT_OK;
  }
  PostWs w = carve(workspace, n_seg, n_samples);
  const int out_first = prm->stitch ? 1 : 0;
  k_plan<<<1, 256, 0, st>>>(w.plan, w.hdr, seg_off, n_seg, prm->normalize, prm->target_peak,
                            peak_override, prm->concat, prm->stitch ? prm->gap_frames : 0,
                            prm->stitch ? prm->stitch_tail : 1, results);
  VT_LAUNCHED();
  if (total_out) {
    k_total_out<<<1, 32, 0, st>>>(w.hdr, total_out);
    VT_LAUNCHED();
  }
  dim3 g(grid_x_for(n_seg, max_seg_len + (prm->stitch ? prm->gap_frames : 0), kWriteTile), n_seg);
  if (prm->out_pcm16)
    k_write<true><<<g, kThreads, 0, st>>>(audio, seg_off, n_seg, w.plan, out_first, prm->clip, out,
                                          out_capacity);
  else
    k_write<false><<<g, kThreads, 0, st>>>(audio, seg_off, n_seg, w.plan, out_first, prm->clip, out,
                                           out_capacity);
  VT_LAUNCHED();
  return VT_OK;
}

int vt_post_process(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples,
                    int64_t max_seg_len, const vt_post_params* prm, void* out, int64_t out_capacity,
                    double* results, int64_t* total_out, void* workspace, int64_t workspace_bytes,
                    void* stream) {
  launch_counter() = 0;
  int rc = vt_post_analyze(audio, seg_off, n_seg, n_samples, max_seg_len, prm, nullptr, workspace,
                           workspace_bytes, stream);
  if (rc) return rc;
  return vt_post_write(audio, seg_off, n_seg, n_samples, max_seg_len, prm, nullptr, out, out_capacity,
                       results, total_out, workspace, workspace_bytes, stream);
}

int vt_find_active_range(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples,
                         int64_t max_seg_len, float threshold, int min_silence_frames, int64_t* ranges,
                         void* workspace, int64_t workspace_bytes, void* stream_v) {
  int rc = check_audio(audio, seg_off, n_seg);
  if (rc) return rc;
  VT_REQUIRE(ranges != nullptr || n_seg == 0, "ranges is NULL");
  VT_REQUIRE(workspace_bytes >= post_ws_bytes(n_seg, n_samples), "post workspace too small");
  if (n_seg == 0) return VT_OK;
  launch_counter() = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  PostWs w = carve(workspace, n_seg, n_samples);
  k_post_init<<<(n_seg + 255) / 256, 256, 0, st>>>(w.plan, w.hdr, seg_off, n_seg, nullptr);
  VT_LAUNCHED();
  k_tile_bases<<<1, 256, 0, st>>>(w.plan, seg_off, n_seg);
  VT_LAUNCHED();
  dim3 g(grid_x_for(n_seg, max_seg_len, kScanTile * (kThreads / 32)), n_seg);
  k_scan<<<g, kThreads, 0, st>>>(audio, seg_off, n_seg, threshold, 1, w.plan, w.tile_max);
  VT_LAUNCHED();
  k_range_out<<<(n_seg + 255) / 256, 256, 0, st>>>(w.plan, seg_off, n_seg, min_silence_frames, ranges);
  VT_LAUNCHED();
  return VT_OK;
}

int vt_post_stats(const float* audio, const int64_t* seg_off, int n_seg, int64_t n_samples, int64_t max_seg_len,
                  float threshold, int64_t* first_last, float* peak, void* workspace, int64_t workspace_bytes,
                  void* stream_v) {
  int rc = check_audio(audio, seg_off, n_seg);
  if (rc) return rc;
  VT_REQUIRE((first_last != nullptr && peak != nullptr) || n_seg == 0, "vt_post_stats: NULL output");
  VT_REQUIRE(workspace_bytes >= post_ws_bytes(n_seg, n_samples), "post workspace too small");
  if (n_seg == 0) return VT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);     // (the launch counter keeps running: part of a job)
  PostWs w = carve(workspace, n_seg, n_samples);
  k_post_init<<<(n_seg + 255) / 256, 256, 0, st>>>(w.plan, w.hdr, seg_off, n_seg, nullptr);
  VT_LAUNCHED();
  k_tile_bases<<<1, 256, 0, st>>>(w.plan, seg_off, n_seg);
  VT_LAUNCHED();
  dim3 g(grid_x_for(n_seg, max_seg_len, kScanTile * (kThreads / 32)), n_seg);
  k_scan<<<g, kThreads, 0, st>>>(audio, seg_off, n_seg, threshold, 1, w.plan, w.tile_max);
  VT_LAUNCHED();
  k_seg_stats<<<n_seg, kThreads, 0, st>>>(w.plan, seg_off, w.tile_max, first_last, peak);
  VT_LAUNCHED();
  return VT_OK;
}

int vt_snap_zero_crossing(const float* audio, const int64_t* seg_off, int n_seg, const int64_t* idx_in,
                          int radius_samples, int64_t* idx_out, void* stream_v) {
  int rc = check_audio(audio, seg_off, n_seg);
  if (rc) return rc;
  VT_REQUIRE(radius_samples >= 0, "radius must be >= 0");
  if (n_seg == 0) return VT_OK;
  launch_counter() = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  k_snap_only<<<n_seg, kThreads, 0, st>>>(audio, seg_off, idx_in, radius_samples, idx_out);
  VT_LAUNCHED();
  return VT_OK;
}

int vt_rms(const float* audio, const int64_t* seg_off, int n_seg, double* rms_out, void* workspace, int64_t workspace_bytes,
           void* stream_v) {
  VT_REQUIRE(n_seg >= 0 && n_seg <= 65535, "vt_rms: n_seg must be in [0, 65535]");
  if (n_seg == 0) return VT_OK;
  VT_REQUIRE(audio && seg_off && rms_out && workspace, "vt_rms: NULL argument");
  VT_REQUIRE(workspace_bytes >= (int64_t)n_seg * VT_RMS_PARTIALS * 8, "vt_rms: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  k_rms_partial<<<dim3(VT_RMS_PARTIALS, n_seg), kThreads, 0, st>>>(audio, seg_off, reinterpret_cast<double*>(workspace));
  VT_LAUNCHED();
  k_rms_final<<<(n_seg + 7) / 8, 256, 0, st>>>(seg_off, n_seg, reinterpret_cast<const double*>(workspace), rms_out);
  VT_LAUNCHED();
  return VT_OK;
}

int vt_wav_pcm16_header(void* dst, int sample_rate, const int64_t* n_samples_dev, int64_t n_samples_host, void* stream_v) {
  VT_REQUIRE(dst != nullptr, "vt_wav_pcm16_header: dst is NULL");
  VT_REQUIRE(sample_rate > 0 && n_samples_host >= 0 && n_samples_host < (1LL << 31) - 64, "vt_wav_pcm16_header: bad sample rate / length");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  k_wav_header<<<1, 32, 0, st>>>(reinterpret_cast<unsigned char*>(dst), sample_rate,
                                 reinterpret_cast<const long long*>(n_samples_dev), (long long)n_samples_host);
  VT_LAUNCHED();
  return VT_OK;
}

int vt_pcm16_encode(const float* in, int16_t* out, int64_t n, void* stream_v) {
  VT_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return VT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  int blocks = (int)((n + 1023) / 1024 < 148 * 8 ? (n + 1023) / 1024 : 148 * 8);
  k_pcm16_encode<<<blocks, 256, 0, st>>>(in, reinterpret_cast<short*>(out), n);
  VT_LAUNCHED();
  return VT_OK;
}

int vt_pcm16_decode(const int16_t* in, float* out, int64_t n, void* stream_v) {
  VT_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return VT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  int blocks = (int)((n + 1023) / 1024 < 148 * 8 ? (n + 1023) / 1024 : 148 * 8);
  k_pcm16_decode<<<blocks, 256, 0, st>>>(reinterpret_cast<const short*>(in), out, n);
  VT_LAUNCHED();
  return VT_OK;
}

}  // extern "C"
