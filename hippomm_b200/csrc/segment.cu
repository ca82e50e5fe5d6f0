// hippo_segment_boundaries: the greedy boundary state machine of _segment_sequence
// (hm:1002-1114), operation for operation in fp64 (this file is compiled with -fmad=false
// so that no product feeds a fused add the Python interpreter would have rounded).
//
// One CTA of 1024 threads per stream.  The chain over segments is sequential (each boundary
// anchors the next window), so the kernel is bound by the latency of ONE segment, which is kept
// at about one global-memory round trip plus three barriers:
//   * frame times and adjacent-pair SSIMs are staged in shared memory once (up to kStageFrames
//     frames);
//   * the window's frame range [lo, hi] (hm:1045-1048) is found by all threads at once, each testing one
//     frame from the previous window's start onwards (boundaries only move forward) -- no binary search;
//   * the backward SSIM scan (hm:1052-1059) tests 1024 pairs per step, the latest hit wins (atomicMax);
//   * the audio scan (hm:1061-1077) evaluates 64 half-second windows per step, one HALF-WARP per window:
//     a 30 s span has at most 59, so one step covers it.  The lanes split the window's pyramid terms
//     (all loads independent, issued together) and reduce with shuffles; the first window in the
//     reference's order that is below the threshold wins (atomicMin).  The video work runs in the
//     shadow of these loads.
// The scalar state (current_start, current_end, optimal_end) is carried redundantly by all threads.
// Window sums are exact for int16-origin PCM (every term is an integer multiple of 2^-30 below 2^53),
// hence independent of the summation order.
#include "audio.cuh"

namespace hippo {

constexpr int kSegThreads = 1024;
constexpr int kSegWindows = kSegThreads / 16;   // audio windows per step
constexpr int kStageFrames = 6000;     // 2 x 6000 doubles = 96 KB of dynamic shared memory

__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }  // Python min(a, b)

// Sum of squares of samples [s, e) by one HALF-WARP (hl = lane & 15): head samples up to a 16-boundary,
// 16-blocks up to a 512-boundary, 512-blocks, 16-blocks, tail samples -- the same terms as
// window_sumsq_pyramid, dealt round-robin to the 16 lanes.  Every lane returns the sum.
__device__ __forceinline__ double halfwarp_window_sumsq(const void* pcm, int dtype, int nch,
                                                        const double* __restrict__ e16,
                                                        const double* __restrict__ e512, int64_t s, int64_t e,
                                                        int hl) {
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
    // edge samples: at most 15 at either end (31 when the window lies inside one 16-block), two per lane
    double ve = 0.0;
    for (int t = hl; t < n_head; t += 16) { const double x = pcm_mono(pcm, dtype, nch, s + t); ve += x * x; }
    for (int t = hl; t < n_tail; t += 16) { const double x = pcm_mono(pcm, dtype, nch, b16 + t); ve += x * x; }
    // pyramid terms, dealt round-robin; the first eight per lane are issued together (independent loads)
    const int total = n_lo16 + n_512 + n_hi16;
    const double* lo16p = e16 + (a16 >> 4);
    const double* midp = e512 + (a512 >> 9) - n_lo16;
    const double* hi16p = e16 + (b512 >> 4) - n_lo16 - n_512;
    double vp[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = hl + 16 * u;
      const double* ptr = t < n_lo16 ? lo16p : (t < n_lo16 + n_512 ? midp : hi16p);
      vp[u] = t < total ? ptr[t] : 0.0;
    }
    acc = ve + ((vp[0] + vp[1]) + (vp[2] + vp[3])) + ((vp[4] + vp[5]) + (vp[6] + vp[7]));
    for (int t = hl + 128; t < total; t += 16) {
      const double* ptr = t < n_lo16 ? lo16p : (t < n_lo16 + n_512 ? midp : hi16p);
      acc += ptr[t];
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // stays inside the half-warp
  return acc;
}

__global__ void __launch_bounds__(kSegThreads, 1) segment_kernel(const hippo_stream_desc* __restrict__ streams,
                                                                 int nstreams, double max_dur, double min_dur,
                                                                 double ssim_thr, double db_thr) {
  extern __shared__ double s_stage[];          // [kStageFrames] frame times, [kStageFrames] ssim
  __shared__ long long s_lo, s_hi, s_vpick;
  __shared__ int s_apick;
  const int tid = threadIdx.x;
  const int si = blockIdx.x;
  if (si >= nstreams) return;
  const hippo_stream_desc S = streams[si];

  const bool has_video = S.frame_times != nullptr && S.nframes > 0;   // `if video_frames and frame_times`
  const bool has_audio = S.pcm != nullptr && S.sample_rate != 0.0;    // `audio_data is not None and audio_sample_rate`
  const double sr = S.sample_rate;
  const int64_t nf = S.nframes;

  const double* ftimes = S.frame_times;
  const double* ssim = S.ssim;
  if (has_video && nf <= kStageFrames) {
    for (int64_t i = tid; i < nf; i += kSegThreads) s_stage[i] = S.frame_times[i];
    ftimes = s_stage;
    if (S.ssim != nullptr) {
      for (int64_t i = tid; i < nf - 1; i += kSegThreads) s_stage[kStageFrames + i] = S.ssim[i];
      ssim = s_stage + kStageFrames;
    }
  }
  __syncthreads();

  // hm:1027-1032
  double total;
  if (has_video) total = ftimes[nf - 1] - ftimes[0];
  else if (has_audio) total = (double)S.ns / sr;
  else { if (tid == 0) *S.out_count = 0; return; }

  const int64_t w = has_audio ? (int64_t)(0.5 * sr) : 0;   // hm:1066; the host rejects w < 1 like range() does
  int count = 0;
  bool overflow = false;
  int64_t hint = 0;                                  // first frame with t >= current_start so far
  double cs = 0.0;                                   // hm:1034
  while (cs < total) {                               // hm:1036
    const double ce = py_min(cs + max_dur, total);   // hm:1038
    double opt = ce;                                 // hm:1041
    const int64_t s0 = (int64_t)(cs * sr);           // int() truncates toward zero
    const int64_t e0 = (int64_t)(ce * sr);
    const int64_t first = e0 - s0 - w;               // hm:1068: range(first, 0, -w)

    if (tid == 0) { s_lo = -1; s_hi = -2; s_vpick = -1; s_apick = 0x7fffffff; }
    __syncthreads();

    // ---- video, step 1: indices with cs <= t <= ce form one run [lo, hi] (frame_times is non-decreasing):
    // lo = first t >= cs (>= hint), hi = last t <= ce (>= hint - 1)
    int64_t lo = nf, hi = nf - 1;
    if (has_video) {
      if (tid == 0 && (hint >= nf || ftimes[hint] > ce)) s_hi = hint - 1;
      for (int64_t base = hint;; base += kSegThreads) {
        const int64_t i = base + tid;
        if (i < nf) {
          const double t = ftimes[i];
          if (t >= cs && (i == hint || ftimes[i - 1] < cs)) s_lo = i;
          if (t <= ce && (i == nf - 1 || ftimes[i + 1] > ce)) s_hi = i;
        }
        if (base + kSegThreads >= nf) break;
        __syncthreads();
        if (s_lo >= 0 && s_hi >= -1) break;          // uniform: read after the barrier, written before it
        __syncthreads();
      }
    }

    // ---- audio (hm:1061-1077; runs second in the reference and overwrites the video boundary): first step
    // of 64 windows issued now, its loads overlap the rest of the video work
    bool more_audio = false;
    if (has_audio && w >= 1) {
      const int64_t i = first - (int64_t)(tid >> 4) * w;   // one half-warp per window
      double ss = 0.0;
      int64_t ws = 0, we = 0;
      if (i > 0) {
        ws = s0 + i; we = ws + w;                    // audio_data[window_start:window_end] clips
        if (ws > S.ns) ws = S.ns;
        if (we > S.ns) we = S.ns;
      }
      ss = halfwarp_window_sumsq(S.pcm, S.pcm_dtype, S.nch, S.e16, S.e512, ws, we, tid & 15);
      if (i > 0 && (tid & 15) == 0 && level_db(ss, we - ws) < db_thr) atomicMin(&s_apick, tid >> 4);
      more_audio = first - (int64_t)kSegWindows * w > 0;
    }
    __syncthreads();

    // ---- video, step 2 (hm:1050-1059): scan i = hi .. lo+1, pair (frame i, frame i-1) = ssim[i-1]
    if (has_video) {
      lo = s_lo >= 0 ? s_lo : nf;
      hi = s_hi >= -1 ? s_hi : nf - 1;
      hint = lo;
    }
    int apick = s_apick;
    if (has_video && hi - lo + 1 > 1 && ssim != nullptr) {
      for (int64_t top = hi; top > lo; top -= kSegThreads) {
        const int64_t i = top - tid;
        if (i > lo && ssim[i - 1] < ssim_thr) atomicMax(&s_vpick, (long long)i);   // NaN compares false, as in Python
        __syncthreads();
        const long long pick = s_vpick;
        if (pick >= 0) { opt = ftimes[pick]; break; }
        if (top - kSegThreads > lo) __syncthreads();
      }
    }

    if (has_audio && w >= 1) {
      // further steps only when the span holds more than 64 windows (max_segment_duration > 32 s)
      int64_t base = first;
      while (apick == 0x7fffffff && more_audio) {
        base -= (int64_t)kSegWindows * w;
        __syncthreads();                              // everyone has read s_apick
        const int64_t i = base - (int64_t)(tid >> 4) * w;
        int64_t ws = 0, we = 0;
        if (i > 0) {
          ws = s0 + i; we = ws + w;
          if (ws > S.ns) ws = S.ns;
          if (we > S.ns) we = S.ns;
        }
        const double ss = halfwarp_window_sumsq(S.pcm, S.pcm_dtype, S.nch, S.e16, S.e512, ws, we, tid & 15);
        if (i > 0 && (tid & 15) == 0 && level_db(ss, we - ws) < db_thr) atomicMin(&s_apick, tid >> 4);
        __syncthreads();
        apick = s_apick;
        more_audio = base - (int64_t)kSegWindows * w > 0;
      }
      if (apick != 0x7fffffff) {
        const int64_t iw = base - (int64_t)apick * w;
        opt = (double)(s0 + iw) / sr;                // hm:1076
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
    __syncthreads();                                 // the shared picks are re-armed at the top
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
