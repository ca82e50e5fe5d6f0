// hippo_segment_boundaries: the greedy boundary state machine of _segment_sequence
// (hm:1002-1114), operation for operation in fp64 (this file is compiled with -fmad=false
// so that no product feeds a fused add the Python interpreter would have rounded).
//
// One warp per stream.  The scalar state (current_start, current_end, optimal_end) is
// carried redundantly by all lanes; lanes only diverge to evaluate up to 32 candidate
// frame pairs / audio windows of one scan at a time, and a ballot picks the first hit in
// the reference's scan order (backwards from the window end).
#include "audio.cuh"

namespace hippo {

__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }  // Python min(a, b)

__global__ void __launch_bounds__(128) segment_kernel(const hippo_stream_desc* __restrict__ streams, int nstreams,
                                                      double max_dur, double min_dur, double ssim_thr,
                                                      double db_thr) {
  const int lane = threadIdx.x & 31;
  const int si = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (si >= nstreams) return;
  const hippo_stream_desc S = streams[si];

  const bool has_video = S.frame_times != nullptr && S.nframes > 0;   // `if video_frames and frame_times`
  const bool has_audio = S.pcm != nullptr && S.sample_rate != 0.0;    // `audio_data is not None and audio_sample_rate`
  const double sr = S.sample_rate;

  // hm:1027-1032
  double total;
  if (has_video) total = S.frame_times[S.nframes - 1] - S.frame_times[0];
  else if (has_audio) total = (double)S.ns / sr;
  else { if (lane == 0) *S.out_count = 0; return; }

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
      while (a < b) { const int64_t m = (a + b) >> 1; if (S.frame_times[m] >= cs) b = m; else a = m + 1; }
      const int64_t lo = a;
      a = 0; b = S.nframes;
      while (a < b) { const int64_t m = (a + b) >> 1; if (S.frame_times[m] <= ce) a = m + 1; else b = m; }
      const int64_t hi = a - 1;
      // hm:1050-1059: scan i = hi .. lo+1, pair (frame i, frame i-1) = ssim[i-1]
      if (hi - lo + 1 > 1 && S.ssim != nullptr) {
        for (int64_t top = hi; top > lo; top -= 32) {
          const int64_t i = top - lane;
          bool hit = false;
          if (i > lo) hit = S.ssim[i - 1] < ssim_thr;        // NaN compares false, as in Python
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (m) { opt = S.frame_times[top - (__ffs(m) - 1)]; break; }
        }
      }
    }

    if (has_audio) {
      // hm:1061-1077 -- runs second and overwrites the video boundary
      const int64_t s0 = (int64_t)(cs * sr);         // int() truncates toward zero
      const int64_t e0 = (int64_t)(ce * sr);
      const int64_t w = (int64_t)(0.5 * sr);
      const int64_t first = e0 - s0 - w;             // range(first, 0, -w)
      for (int64_t base = first; base > 0; base -= 32 * w) {
        const int64_t i = base - (int64_t)lane * w;
        bool hit = false;
        if (i > 0) {
          int64_t ws = s0 + i, we = ws + w;          // audio_data[window_start:window_end] clips
          if (ws > S.ns) ws = S.ns;
          if (we > S.ns) we = S.ns;
          double ss = 0.0;
          if (we > ws) ss = window_sumsq_pyramid(S.pcm, S.pcm_dtype, S.nch, S.e16, S.e512, ws, we);
          hit = level_db(ss, we - ws) < db_thr;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
          const int64_t iw = base - (int64_t)(__ffs(m) - 1) * w;
          opt = (double)(s0 + iw) / sr;              // hm:1076
          break;
        }
      }
    }

    // hm:1080-1084
    if (opt - cs < min_dur) opt = py_min(cs + min_dur, total);

    if (count < S.max_segments) {
      if (lane == 0) { S.out_bounds[2 * count] = cs; S.out_bounds[2 * count + 1] = opt; }
    } else {
      overflow = true;
      break;
    }
    ++count;
    cs = opt;                                        // hm:1111
  }
  if (lane == 0) *S.out_count = overflow ? -1 : count;
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
  segment_kernel<<<(nstreams + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      streams, nstreams, max_segment_duration, min_segment_duration, frame_similarity_threshold,
      audio_silence_threshold);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}
