// Rational-ratio polyphase resampler: the step between synthesize_chunk and stitching for engines whose rate differs from
// the pipeline's 24 kHz (reference backend/shared/tts_pipeline.py:100-111 -> librosa.resample, called at :389-390; e.g.
// CosyVoice at 22 050 Hz, tts_backends/cosyvoice_runner.py:84,131).  Segment-batched like the rest of the post path.
//
//   y[m] = sum_j H[(m*down) % up][j] * x[(m*down) / up + J0 - j],   j in [0, 2*J0],  x = 0 outside the segment
//
// H is the phase table of a Kaiser-windowed sinc built on the host in float64 (vocalie-tts_b200/post.py, same numbers as
// oracle/resample_oracle.py) and rounded to fp32.  HBM-bound byte work: 4 B read per input sample + 4 B written per output
// sample; the table (up x (2*J0+1) floats: 83 KB for 22 050 -> 24 000) and the overlapping input windows live in L1/L2.
// A block produces kOutTile consecutive outputs of one segment from an input span staged in shared memory with coalesced
// loads; a thread's outputs are kThreads apart, so neighbouring threads read neighbouring shared-memory words (phases of
// consecutive outputs differ, so the table rows are read through the read-only cache).
#include "vt_common.cuh"

namespace vt {

constexpr int kRsThreads = 256;
constexpr int kRsOutTile = 2048;           // outputs per block iteration
constexpr int kRsMaxSpan = 8192;           // staged input samples per tile (floats): covers down/up <= 3.9 with 128 taps

__global__ void __launch_bounds__(kRsThreads)
k_resample(const float* __restrict__ in, const int64_t* __restrict__ off_in, const int64_t* __restrict__ off_out,
           const float* __restrict__ table, int up, int down, int ntaps, float* __restrict__ out) {
  __shared__ float sx[kRsMaxSpan];
  const int seg = blockIdx.y;
  const long long a = off_in[seg], n_in = off_in[seg + 1] - a;
  const long long o0 = off_out[seg], n_out = off_out[seg + 1] - o0;
  const int j0 = (ntaps - 1) / 2;
  const long long n_tiles = (n_out + kRsOutTile - 1) / kRsOutTile;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long m0 = tile * kRsOutTile;
    const long long m1 = m0 + kRsOutTile < n_out ? m0 + kRsOutTile : n_out;
    // input indices touched by outputs [m0, m1): i0(m) + j0 - j for j in [0, ntaps)
    const long long lo = (m0 * down) / up - j0;
    const long long hi = ((m1 - 1) * down) / up + j0;           // inclusive
    const int span = (int)(hi - lo + 1);
    __syncthreads();
    for (int i = threadIdx.x; i < span; i += kRsThreads) {
      const long long k = lo + i;
      sx[i] = (k >= 0 && k < n_in) ? __ldg(in + a + k) : 0.0f;
    }
    __syncthreads();
    for (long long m = m0 + threadIdx.x; m < m1; m += kRsThreads) {
      const long long q = m * down;
      const int p = (int)(q % up);
      const int base = (int)(q / up + j0 - lo);                 // shared index of x[i0 + j0]
      const float* __restrict__ h = table + (size_t)p * ntaps;
      float acc0 = 0.0f, acc1 = 0.0f;
      int j = 0;
      for (; j + 1 < ntaps; j += 2) {
        acc0 = fmaf(__ldg(h + j), sx[base - j], acc0);
        acc1 = fmaf(__ldg(h + j + 1), sx[base - j - 1], acc1);
      }
      if (j < ntaps) acc0 = fmaf(__ldg(h + j), sx[base - j], acc0);
      out[o0 + m] = acc0 + acc1;
    }
  }
}

}  // namespace vt

using namespace vt;

extern "C" int vt_resample(const float* in, const int64_t* seg_off_in, const int64_t* seg_off_out, int n_seg,
                           int64_t max_out_len, int up, int down, const float* table, int ntaps, float* out,
                           void* stream_v) {
  VT_REQUIRE(n_seg >= 0 && n_seg <= 65535, "vt_resample: n_seg must be in [0, 65535]");
  if (n_seg == 0) return VT_OK;
  VT_REQUIRE(in && seg_off_in && seg_off_out && table && out, "vt_resample: NULL argument");
  VT_REQUIRE(up >= 1 && down >= 1 && ntaps >= 1 && (ntaps & 1), "vt_resample: bad ratio / tap count");
  // the staged span of one tile: kRsOutTile * down / up + ntaps samples
  VT_REQUIRE((long long)kRsOutTile * down / up + ntaps + 2 <= kRsMaxSpan,
             "vt_resample: ratio %d/%d with %d taps per phase exceeds the staged window (down/up too large)", up, down, ntaps);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  long long tiles = (max_out_len + kRsOutTile - 1) / kRsOutTile;
  long long cap = (148LL * 8 + n_seg - 1) / n_seg;
  if (tiles > cap) tiles = cap;
  if (tiles < 1) tiles = 1;
  k_resample<<<dim3((unsigned)tiles, (unsigned)n_seg), kRsThreads, 0, st>>>(in, seg_off_in, seg_off_out, table, up, down, ntaps, out);
  VT_LAUNCHED();
  return VT_OK;
}
