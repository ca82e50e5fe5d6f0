// hippo_segment_boundaries: the greedy boundary state machine of _segment_sequence
// (hm:1002-1114), operation for operation in fp64 (this file is compiled with -fmad=false
// so that no product feeds a fused add the Python interpreter would have rounded).
//
// One CTA of 1024 threads per stream.  The chain over segments is sequential (each boundary
// anchors the next window), so the kernel is bound by the latency of one segment's scan:
//   * frame times and adjacent-pair SSIMs are staged in shared memory once (up to kStageFrames
//     frames), so the two binary searches and the backward SSIM scan never touch global memory;
//   * the backward scan tests 1024 frame pairs per step, a ballot picks the first hit in the
//     reference's order;
//   * the audio scan evaluates 32 half-second windows per step, one warp per window: the lanes split
//     the window's pyramid terms (<= 107 loads, all independent) and reduce with shuffles, so a
//     step costs about one memory latency instead of a chain of ~100.
// The scalar state (current_start, current_end, optimal_end) is carried redundantly by all threads.
// Window sums are exact for int16-origin PCM (every term is an integer multiple of 2^-30 below 2^53),
// hence independent of the summation order.
#include "audio.cuh"

namespace hippo {

constexpr int kSegThreads = 1024;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kStageFrames = 6000;     // 2 x 6000 doubles = 96 KB of dynamic shared memory

__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }  // Python min(a, b)

// Sum of squares of samples [s, e) by one WARP: head samples up to a 16-boundary, 16-blocks up to a
// 512-boundary, 512-blocks, 16-blocks, tail samples -- the same terms as window_sumsq_pyramid, dealt
// round-robin to the lanes.
__device__ __forceinline__ double warp_window_sumsq(const void* pcm, int dtype, int nch,
                                                    const double* __restrict__ e16,
                                                    const double* __restrict__ e512, int64_t s, int64_t e,
                                                    int lane) {
  double acc = 0.0;
  if (e > s) {
    int64_t a16 = (s + 15) & ~(int64_t)15;      // first 16-boundary >= s
    int64_t b16 = e & ~(int64_t)15;             // last 16-boundary <= e
    if (a16 > b16) { a16 = e; b16 = e; }        // window inside one 16-block: samples only
    int64_t a512 = (a16 + 511) & ~(int64_t)511;
    int64_t b512 = b16 & ~(int64_t)511;
    if (a512 > b512) { a512 = b16; b512 = b16; }  // no whole 512-block: 16-blocks only
    const int n_head = (int)(a16 - s);
    const int n_lo16 = (int)((a512 - a16) >> 4);
    const int n_512 = (int)((b512 - a512) >> 9);
    const int n_hi16 = (int)((b16 - b512) >> 4);
    const int n_tail = (int)(e - b16);
    // edge samples: lanes 0-14 take the head, lanes 16-30 the tail (at most 15 each), one load per lane
    double ve = 0.0;
    if (lane < 15) { if (lane < n_head) { const double x = pcm_mono(pcm, dtype, nch, s + lane); ve = x * x; } }
    else if (lane >= 16 && lane - 16 < n_tail) { const double x = pcm_mono(pcm, dtype, nch, b16 + (lane - 16)); ve = x * x; }
    // pyramid terms, dealt round-robin; the first three per lane are issued together (independent loads)
    const int total = n_lo16 + n_512 + n_hi16;
    const double* lo16p = e16 + (a16 >> 4);
    const double* midp = e512 + (a512 >> 9) - n_lo16;
    const double* hi16p = e16 + (b512 >> 4) - n_lo16 - n_512;
    double vp[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int t = lane + 32 * u;
      const double* ptr = t < n_lo16 ? lo16p : (t < n_lo16 + n_512 ? midp : hi16p);
      vp[u] = t < total ? ptr[t] : 0.0;
    }
    acc = ve + vp[0] + vp[1] + vp[2];
    for (int t = lane + 96; t < total; t += 32) {
      const double* ptr = t < n_lo16 ? lo16p : (t < n_lo16 + n_512 ? midp : hi16p);
      acc += ptr[t];
    }
  }
  return warp_sum(acc);
}

__global__ void __launch_bounds__(kSegThreads, 1) segment_kernel(const hippo_stream_desc* __restrict__ streams,
                                                                 int nstreams, double max_dur, double min_dur,
                                                                 double ssim_thr, double db_thr) {
  extern __shared__ double s_stage[];          // [kStageFrames] frame times, [kStageFrames] ssim
  __shared__ int s_pick[kSegWarps];
  __shared__ int s_first;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int si = blockIdx.x;
  if (si >= nstreams) return;
  const hippo_stream_desc S = streams[si];

  const bool has_video = S.frame_times != nullptr && S.nframes > 0;   // `if video_frames and frame_times`
  const bool has_audio = S.pcm != nullptr && S.sample_rate != 0.0;    // `audio_data is not None and audio_sample_rate`
  const double sr = S.sample_rate;

  const double* ftimes = S.frame_times;
  const double* ssim = S.ssim;
  if (has_video && S.nframes <= kStageFrames) {
    for (int64_t i = tid; i < S.nframes; i += kSegThreads) s_stage[i] = S.frame_times[i];
    ftimes = s_stage;
    if (S.ssim != nullptr) {
      for (int64_t i = tid; i < S.nframes - 1; i += kSegThreads) s_stage[kStageFrames + i] = S.ssim[i];
      ssim = s_stage + kStageFrames;
    }
  }
  __syncthreads();

  // hm:1027-1032
  double total;
  if (has_video) total = ftimes[S.nframes - 1] - ftimes[0];
  else if (has_audio) total = (double)S.ns / sr;
  else { if (tid == 0) *S.out_count = 0; return; }

  int count = 0;
  bool overflow = false;
  double cs = 0.0;                                   // hm:1034
  while (cs < total) {                               // hm:1036
    const double ce = py_min(cs + max_dur, total);   // hm:1038
    double opt = ce;                                 // hm:1041

    if (has_video) {
      // hm:1045-1048: indices with cs <= t <= ce; frame_times is non-decreasing, so they form
      // one run [lo, hi].  lo = first t >= cs, hi = last t <= ce.
      int64_t a = 0, b = S.nframes;
      while (a < b) { const int64_t m = (a + b) >> 1; if (ftimes[m] >= cs) b = m; else a = m + 1; }
      const int64_t lo = a;
      a = 0; b = S.nframes;
      while (a < b) { const int64_t m = (a + b) >> 1; if (ftimes[m] <= ce) a = m + 1; else b = m; }
      const int64_t hi = a - 1;
      // hm:1050-1059: scan i = hi .. lo+1, pair (frame i, frame i-1) = ssim[i-1]
      if (hi - lo + 1 > 1 && ssim != nullptr) {
        for (int64_t top = hi; top > lo; top -= kSegThreads) {
          const int64_t i = top - tid;
          bool hit = false;
          if (i > lo) hit = ssim[i - 1] < ssim_thr;          // NaN compares false, as in Python
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (lane == 0) s_pick[warp] = m ? (warp * 32 + __ffs(m) - 1) : -1;
          __syncthreads();
          int pick = -1;
          for (int w = 0; w < kSegWarps; ++w) { const int v = s_pick[w]; if (v >= 0) { pick = v; break; } }
          __syncthreads();
          if (pick >= 0) { opt = ftimes[top - pick]; break; }
        }
      }
    }

    if (has_audio) {
      // hm:1061-1077 -- runs second and overwrites the video boundary
      const int64_t s0 = (int64_t)(cs * sr);         // int() truncates toward zero
      const int64_t e0 = (int64_t)(ce * sr);
      const int64_t w = (int64_t)(0.5 * sr);
      const int64_t first = e0 - s0 - w;             // range(first, 0, -w)
      for (int64_t base = first; base > 0; base -= kSegWarps * w) {
        if (tid == 0) s_first = kSegWarps;
        __syncthreads();
        const int64_t i = base - (int64_t)warp * w;  // one warp per window
        if (i > 0) {
          int64_t ws = s0 + i, we = ws + w;          // audio_data[window_start:window_end] clips
          if (ws > S.ns) ws = S.ns;
          if (we > S.ns) we = S.ns;
          const double ss = warp_window_sumsq(S.pcm, S.pcm_dtype, S.nch, S.e16, S.e512, ws, we, lane);
          if (lane == 0 && level_db(ss, we - ws) < db_thr) atomicMin(&s_first, warp);
        }
        __syncthreads();
        const int pick = s_first;
        __syncthreads();
        if (pick < kSegWarps) {
          const int64_t iw = base - (int64_t)pick * w;
          opt = (double)(s0 + iw) / sr;              // hm:1076
          break;
        }
      }
    }

    // hm:1080-1084
    if (opt - cs < min_dur) opt = py_min(cs + min_dur, total);

    if (count < S.max_segments) {
      if (tid == 0) { S.out_bounds[2 * count] = cs; S.out_bounds[2 * count + 1] = opt; }
    } else {
      overflow = true;
      break;
    }
    ++count;
    cs = opt;                                        // hm:1111
  }
  if (tid == 0) *S.out_count = overflow ? -1 : count;
}

}  // namespace hippo

extern "C" hippo_status hippo_segment_boundaries(const hippo_stream_desc* streams, int32_t nstreams,
                                                 double max_segment_duration, double min_segment_duration,
                                                 double frame_similarity_threshold,
                                                 double audio_silence_threshold, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(nstreams >= 0, "hippo_segment_boundaries: nstreams < 0");
  if (nstreams == 0) return HIPPO_OK;
  HIPPO_REQUIRE(streams != nullptr, "hippo_segment_boundaries: null stream table");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  const size_t smem = (size_t)2 * kStageFrames * sizeof(double);
  HIPPO_CUDA(cudaFuncSetAttribute(segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  segment_kernel<<<nstreams, kSegThreads, smem, (cudaStream_t)stream>>>(
      streams, nstreams, max_segment_duration, min_segment_duration, frame_similarity_threshold,
      audio_silence_threshold);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}
