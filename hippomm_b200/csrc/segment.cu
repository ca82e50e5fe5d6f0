// hippo_segment_boundaries: the greedy boundary state machine of _segment_sequence
// (hm:1002-1114), operation for operation in fp64 (this file is compiled with -fmad=false
// so that no product feeds a fused add the Python interpreter would have rounded).
//
// One CTA of 256 threads per stream.  The chain over segments is sequential (each boundary
// anchors the next window), so the kernel is bound by the latency of ONE segment -- the dependent
// instruction chain of one thread plus one global-memory round trip and ONE barrier:
//   * frame times and adjacent-pair SSIMs are staged in shared memory once (up to kStageFrames
//     frames);
//   * the window's frame range and the backward SSIM scan (hm:1045-1059) are one pass: every thread tests
//     one frame from the previous window's start onwards (boundaries only move forward, no binary search),
//     membership in [lo+1, hi] is decided from the frame's own and its predecessor's time, the latest pair
//     below the threshold wins (atomicMax);
//   * the audio scan (hm:1061-1077) looks at up to 64 half-second windows per step (a 30 s span has 59), ONE
//     thread per window on warps 0-1 while warps 2-7 run the video scan.  A window is first classified from the
//     512-sample level of the energy pyramid alone: the blocks that lie fully inside it bound its sum of squares
//     from below, the blocks that cover it from above (~17 contiguous loads, no edge samples); a window whose
//     bounds fall on the same side of the power threshold is decided.  Only when the EARLIEST candidate of a
//     step is undecided (its mean within about half a dB of the threshold) the step is redone exactly: four
//     lanes per window split the window's pyramid terms and edge samples (all loads independent, issued
//     together) and reduce with shuffles.  The first window in the reference's order that is below the
//     threshold wins (atomicMin).  The threshold test runs in the power domain (no sqrt / log10 unless the
//     mean is within 1e-12 of the threshold);
//   * the per-segment picks are triple-buffered in shared memory, so video, audio and the re-arming of
//     the next segment's slots need no barrier between them.
// The scalar state (current_start, current_end, optimal_end) is carried redundantly by all threads, which
// is why the CTA is small: with 32 warps the redundant scalar code alone cost ~8k issue cycles per segment.
// Window sums are exact for int16-origin PCM (every term is an integer multiple of 2^-30 below 2^53),
// hence independent of the summation order.
#include "audio.cuh"
#include <cstdio>
#include <cstdlib>

namespace hippo {

constexpr int kSegThreads = 256;                // every thread runs the scalar chain redundantly: few warps keep it cheap
constexpr int kSegLanes = 4;                    // lanes per audio window
constexpr int kSegWindows = kSegThreads / kSegLanes;   // 64 audio windows per step
constexpr int kVidThreads = kSegThreads - kSegWindows; // threads 64..255 scan the frames while 0..63 classify windows
constexpr int kSegBatch = 24;                   // pyramid terms a lane loads before it starts adding
constexpr int kStageFrames = 6000;     // 2 x 6000 doubles = 96 KB of dynamic shared memory

__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }  // Python min(a, b)

// Sum of squares of samples [s, e) by kSegLanes adjacent lanes (hl = lane & 3): head samples up to a
// 16-boundary, 16-blocks up to a 512-boundary, 512-blocks, 16-blocks, tail samples -- the same terms as
// window_sumsq_pyramid, dealt round-robin to the lanes.  Every lane of the group returns the sum.
__device__ __forceinline__ double group_window_sumsq(const void* pcm, int dtype, int nch,
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
    // edge samples: at most 15 at either end (31 when the window lies inside one 16-block)
    for (int t = hl; t < n_head; t += kSegLanes) { const double x = pcm_mono(pcm, dtype, nch, s + t); acc += x * x; }
    for (int t = hl; t < n_tail; t += kSegLanes) { const double x = pcm_mono(pcm, dtype, nch, b16 + t); acc += x * x; }
    // pyramid terms, dealt round-robin; 24 per lane are issued together (independent loads): a half-second
    // window at 16 kHz has at most 77 terms, so its whole sum costs one memory round trip
    const int total = n_lo16 + n_512 + n_hi16;
    const double* lo16p = e16 + (a16 >> 4);
    const double* midp = e512 + (a512 >> 9) - n_lo16;
    const double* hi16p = e16 + (b512 >> 4) - n_lo16 - n_512;
    for (int t0 = hl; t0 < total; t0 += kSegBatch * kSegLanes) {
      double vp[kSegBatch];
#pragma unroll
      for (int u = 0; u < kSegBatch; ++u) {
        const int t = t0 + kSegLanes * u;
        const double* ptr = t < n_lo16 ? lo16p : (t < n_lo16 + n_512 ? midp : hi16p);
        vp[u] = t < total ? ptr[t] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kSegBatch; u += 4) acc += (vp[u] + vp[u + 1]) + (vp[u + 2] + vp[u + 3]);
    }
  }
#pragma unroll
  for (int o = kSegLanes / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // stays inside the group
  return acc;
}

// `level_db(sumsq, len) < db_thr` (hm:998-999, hm:1073) without the sqrt / log10 / division on the common path:
// 20 log10(sqrt(m)) < T  <=>  m < 10^(T/10) mathematically, and the reference's fp64 chain is within a few
// ulp of that, so only a mean within 1e-12 (relative) of the power threshold needs the exact evaluation.
__device__ __forceinline__ bool below_level(double sumsq, int64_t len, double db_thr, double pow_thr) {
  if (len <= 0 || !(sumsq > 0.0)) return level_db(sumsq, len) < db_thr;    // -100 (or NaN input): exact path
  if (pow_thr > 0.0 && pow_thr < 1e300) {
    const double bound = pow_thr * (double)len;
    if (sumsq < bound * (1.0 - 1e-12)) return true;
    if (sumsq > bound * (1.0 + 1e-12)) return false;
  }
  return level_db(sumsq, len) < db_thr;
}

// Classification of the window [s, e) (already clipped to the stream) from the 512-sample sums alone:
// 0 = not below the threshold, 1 = below, 2 = undecided.  lower = blocks fully inside the window <= sum of squares
// <= upper = blocks covering it; the last block of the stream holds the samples that exist, so it counts as inside
// when the window ends with the stream.  The 1e-9 margins absorb the rounding of the block sums for float PCM
// (int16-origin sums are exact) and keep the answer identical to below_level() of the exact sum, whose own
// undecided band is 1e-12; an all-zero window (level -100) is "below" only for thresholds above -100 dB.
__device__ __forceinline__ int window_class(const double* __restrict__ e512, int64_t s, int64_t e, int64_t ns,
                                            int64_t w, double bound_w_lo, double bound_w_hi, double db_thr,
                                            double pow_thr) {
  const int64_t len = e - s;
  if (!(pow_thr > 0.0 && pow_thr < 1e300) || !(db_thr > -100.0)) return 2;
  if (len <= 0) return 1;                                         // empty slice: level -100 (NumPy's mean is NaN)
  const int64_t ba = s >> 9, bb = (e - 1) >> 9;                   // covering blocks [ba, bb]
  const bool head_in = (s & 511) == 0, tail_in = (e & 511) == 0 || e == ns;
  // every load below is independent of the others (batches of 16 inner blocks, predicated, immediate offsets), so
  // a half-second window at 16 kHz (at most 17 covering blocks) costs ONE memory round trip
  const double* p = e512 + ba;
  const int nin = (int)(bb - ba) - 1;                             // inner blocks ba+1 .. bb-1
  const double head = p[0];
  const double tail = p[nin + 1 > 0 ? nin + 1 : 0];
  double inner = 0.0;
  for (int u0 = 0; u0 < nin; u0 += 16) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = (u0 + u < nin) ? p[1 + u0 + u] : 0.0;
    inner += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7])) +
             (((v[8] + v[9]) + (v[10] + v[11])) + ((v[12] + v[13]) + (v[14] + v[15])));
  }
  double lower, upper;
  if (nin < 0) {                                                  // one covering block
    lower = (head_in && tail_in) ? head : 0.0;
    upper = head;
  } else {
    lower = inner + (head_in ? head : 0.0) + (tail_in ? tail : 0.0);
    upper = inner + head + tail;
  }
  double lo = bound_w_lo, hi = bound_w_hi;                        // pow_thr * w * (1 -+ 1e-9), hoisted by the caller
  if (len != w) { const double bound = pow_thr * (double)len; lo = bound * (1.0 - 1e-9); hi = bound * (1.0 + 1e-9); }
  if (lower > hi) return 0;
  if (upper < lo) return 1;
  return 2;       // includes NaN sums
}

// Resumable form (states != nullptr): the chain is carried across launches in `states[si]`, and a launch that is not
// the final one only takes the segments whose window [current_start, current_end] is covered by the first
// `frames_ready` frames (their adjacent-pair SSIMs are final) -- so the boundary chain of a stream can run under the
// SSIM kernels of its later frames, one launch per chunk of frames, with no cross-kernel waiting anywhere.
//
// Follow mode (follow_ns > 0, pattern.cu): ONE launch that runs beside the SSIM kernels of the same stream, on an SM
// of its own (the launch asks for most of an SM's shared memory and the SSIM CTAs for a few KB each, so they are never
// placed together: the chain keeps its stand-alone speed instead of sharing issue slots).  The SSIM array starts out
// filled with kSsimPending; a video thread whose pair is still pending polls the pair's 8-byte result (stored once by
// the SSIM warp that completes the pair: the data is its own flag) and keeps the value in shared memory.  Only the
// pairs the reference's scan would look at are ever waited for.  A wait longer than follow_ns -- the SSIM kernels are
// not running beside this one: serialising profiler, sanitizer, launch-blocking debug runs -- suspends the chain
// exactly like a gated pass; the final resumable pass the host always enqueues behind the SSIM kernels completes it.
__global__ void __launch_bounds__(kSegThreads, 1) segment_kernel(const hippo_stream_desc* __restrict__ streams,
                                                                 int nstreams, double max_dur, double min_dur,
                                                                 double ssim_thr, double db_thr,
                                                                 hippo_segment_state* __restrict__ states,
                                                                 int64_t frames_ready, int final_pass,
                                                                 long long follow_ns,
                                                                 const unsigned int* __restrict__ follow_gate,
                                                                 unsigned long long* dbg) {
  extern __shared__ double s_stage[];          // [kStageFrames] frame times, [kStageFrames] ssim
  // per-segment results, triple-buffered so that ONE barrier per segment suffices: segment k uses set k % 3,
  // the last thread re-arms set (k + 1) % 3 at the start of segment k (last read in segment k - 2, which every thread
  // left before the barrier of segment k - 1; first written in segment k + 1, after the barrier of segment k)
  __shared__ long long s_lo[3];
  __shared__ int s_vpick[3];                   // latest pair below the threshold, as frame number - hint (-1 = none)
  __shared__ int s_apick[3], s_amb[3];         // earliest window known to be below the threshold / earliest undecided one
  __shared__ int s_abort;                      // follow mode: a wait timed out
  const int tid = threadIdx.x;
  const int si = blockIdx.x;
  if (si >= nstreams) return;
  const hippo_stream_desc S = streams[si];

  // follow mode: tell the host-side launch order that this CTA is resident (pattern.cu holds the frame kernels back
  // until then: once they fill every SM, no SM would ever drain for a CTA that wants one to itself)
  if (follow_ns > 0 && follow_gate != nullptr && tid == 0) st_volatile_u32(const_cast<unsigned int*>(follow_gate) + 1, 1u);
  if (states != nullptr && states[si].done) return;
  const bool has_video = S.frame_times != nullptr && S.nframes > 0;   // `if video_frames and frame_times`
  const bool has_audio = S.pcm != nullptr && S.sample_rate != 0.0;    // `audio_data is not None and audio_sample_rate`
  const double sr = S.sample_rate;
  const int64_t nf = S.nframes;
  const bool follow = follow_ns > 0 && states != nullptr && has_video;
  const bool staged = has_video && nf <= kStageFrames;    // longer streams: times and SSIMs straight from global memory

  const double* ftimes = S.frame_times;
  const double* ssim = S.ssim;
  if (has_video && nf <= kStageFrames) {
    // eight independent loads in flight per thread: the staging is a handful of memory round trips, not nf / 256
    auto stage = [&](const double* __restrict__ src, double* dst, int64_t cnt) {
      for (int64_t i0 = tid; i0 < cnt; i0 += 8 * kSegThreads) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = i0 + u * kSegThreads < cnt ? src[i0 + u * kSegThreads] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (i0 + u * kSegThreads < cnt) dst[i0 + u * kSegThreads] = v[u];
      }
    };
    stage(S.frame_times, s_stage, nf);
    ftimes = s_stage;
    if (S.ssim != nullptr && follow) {
      // whatever is final by now; the rest stays kSsimPending and is polled by the thread that needs it
      for (int64_t i = tid; i < nf - 1; i += kSegThreads) s_stage[kStageFrames + i] = ld_volatile_f64(S.ssim + i);
      ssim = s_stage + kStageFrames;
    } else if (S.ssim != nullptr) {
      const int64_t pairs = (states != nullptr && !final_pass) ? (frames_ready - 1 < nf - 1 ? frames_ready - 1 : nf - 1) : nf - 1;
      stage(S.ssim, s_stage + kStageFrames, pairs > 0 ? pairs : 0);
      ssim = s_stage + kStageFrames;
    }
  }
  if (tid < 3) { s_lo[tid] = -1; s_vpick[tid] = -1; s_apick[tid] = 0x7fffffff; s_amb[tid] = 0x7fffffff; }
  if (tid == 0) s_abort = 0;
  __syncthreads();

  // hm:1027-1032
  double total;
  if (has_video) total = ftimes[nf - 1] - ftimes[0];
  else if (has_audio) total = (double)S.ns / sr;
  else {
    if (tid == 0) { *S.out_count = 0; if (states != nullptr) states[si].done = 1; }
    return;
  }

  if (follow) {
    // The follower is launched FIRST (an idle GPU gives it its SM at once); nothing it reads is final yet.  Gate: the
    // audio pyramid (a flag stored by a launch behind it) and the first pair's SSIM.
    if (tid == 0) {
      const unsigned long long t0 = globaltimer_ns();
      for (;;) {
        const bool audio_ok = follow_gate == nullptr || !has_audio || ld_volatile_u32(follow_gate) != 0;
        // with audio the chain starts at once: most segments are settled by a silent window and never wait for a pair
        const bool video_ok = S.ssim == nullptr || nf < 2 || (has_audio && follow_gate != nullptr) ||
                              __double_as_longlong(ld_volatile_f64(S.ssim)) != (long long)kSsimPending;
        if (audio_ok && video_ok) break;
        if (globaltimer_ns() - t0 > (unsigned long long)follow_ns) { s_abort = 1; break; }
        __nanosleep(200);
      }
    }
    __syncthreads();
    if (s_abort) return;                       // state untouched: the final pass starts from it
    __threadfence();
  }

  const double pow_thr = pow(10.0, db_thr / 10.0);
  const int64_t w = has_audio ? (int64_t)(0.5 * sr) : 0;   // hm:1066; the host rejects w < 1 like range() does
  const bool scan_audio = has_audio && w >= 1;
  const int vt = tid - kSegWindows;                  // video lane: threads 64..255
  const double bound_w_lo = pow_thr * (double)w * (1.0 - 1e-9), bound_w_hi = pow_thr * (double)w * (1.0 + 1e-9);
  int count = 0;
  bool overflow = false, suspended = false;
  int64_t hint = 0;                                  // first frame with t >= current_start so far
  double cs = 0.0;                                   // hm:1034
  if (states != nullptr) { cs = states[si].current_start; hint = states[si].hint; count = states[si].count; }
  const int count0 = count;
  // a pass that is not final may only look at frames below frames_ready; it needs one of them BEYOND the window
  const bool gated = states != nullptr && !final_pass && has_video && !follow;
  const double t_ready = (gated && frames_ready > 0) ? ftimes[(frames_ready < nf ? frames_ready : nf) - 1] : 0.0;
  long long t_pre = 0, t_bar = 0, t_tail = 0, n_exact = 0;
  while (cs < total) {                               // hm:1036
    const long long c0 = dbg ? clock64() : 0;
    if (gated && !(frames_ready > 0 && t_ready > py_min(cs + max_dur, total))) { suspended = true; break; }
    const int set = (count - count0) % 3;
    if (tid == kSegThreads - 1) {                    // a thread of the video group: off the audio threads' critical path
      const int nx = (count - count0 + 1) % 3;
      s_lo[nx] = -1; s_vpick[nx] = -1; s_apick[nx] = 0x7fffffff; s_amb[nx] = 0x7fffffff;
    }
    const double ce = py_min(cs + max_dur, total);   // hm:1038
    double opt = ce;                                 // hm:1041
    const int64_t s0 = (int64_t)(cs * sr);           // int() truncates toward zero
    const int64_t e0 = (int64_t)(ce * sr);
    const int64_t first = e0 - s0 - w;               // hm:1068: range(first, 0, -w)

    if (tid >= kSegWindows) {
      // ---- video (hm:1045-1059), threads 64..255.  The indices with cs <= t <= ce form one run [lo, hi]
      // (frame_times is non-decreasing) and lo >= hint (boundaries only move forward).  The scan i = hi .. lo+1
      // stops at the first pair (frame i, frame i-1) with ssim[i-1] < threshold, i.e. the LARGEST such i: every
      // thread tests one frame per step -- i is in the scan iff t[i-1] >= cs (i-1 >= lo) and t[i] <= ce (i <= hi)
      // -- and atomicMax keeps it.
      if (has_video) {
        for (int64_t base = hint; base < nf; base += kVidThreads) {
          const int64_t i = base + vt;
          int cand = -1;
          if (i < nf) {
            const double t = ftimes[i];
            const double tp = i > hint ? ftimes[i - 1] : -INFINITY;
            if (t >= cs && !(tp >= cs)) s_lo[set] = i;                       // the unique first frame of the run
            if (ssim != nullptr && i > hint && tp >= cs && t <= ce) {
              double sv = (follow && !staged) ? ld_volatile_f64(S.ssim + (i - 1)) : ssim[i - 1];
              if (follow && __double_as_longlong(sv) == (long long)kSsimPending) {
                // the pair's SSIM warp has not delivered yet: poll its result, keep it for the later segments
                const unsigned long long t0 = globaltimer_ns();
                for (;;) {
                  sv = ld_volatile_f64(S.ssim + (i - 1));
                  if (__double_as_longlong(sv) != (long long)kSsimPending) { if (staged) s_stage[kStageFrames + i - 1] = sv; break; }
                  // a window already KNOWN to be silent settles the segment: the audio boundary overwrites the video
                  // boundary (hm:1061-1077 run second), so this pair's value cannot matter any more -- stop waiting
                  // for it (sv stays pending = NaN: no candidate; the pair is polled again if a later segment needs it)
                  if (scan_audio && *(volatile int*)&s_apick[set] != 0x7fffffff) break;
                  if (globaltimer_ns() - t0 > (unsigned long long)follow_ns) { s_abort = 1; break; }
                  __nanosleep(40);
                }
              }
              if (sv < ssim_thr)              // NaN compares false, as in Python
                cand = (int)(i - hint);       // a run of more than 2^31 frames inside one window does not exist
            }
          }
          const int wmax = __reduce_max_sync(0xffffffffu, cand);             // one shared-memory atomic per warp
          if ((tid & 31) == 0 && wmax >= 0) atomicMax(&s_vpick[set], wmax);
          const int64_t last = base + kVidThreads - 1;
          if (last >= nf - 1 || ftimes[last] > ce) break;                    // uniform: the run ends inside this step
        }
      }
    } else if (scan_audio) {
      // ---- audio (hm:1061-1077; runs second in the reference and overwrites the video boundary), threads 0..63:
      // one window per thread, classified from the 512-sample sums
      const int64_t i = first - (int64_t)tid * w;
      if (i > 0) {
        int64_t ws = s0 + i, we = ws + w;            // audio_data[window_start:window_end] clips
        if (ws > S.ns) ws = S.ns;
        if (we > S.ns) we = S.ns;
        const int cls = window_class(S.e512, ws, we, S.ns, w, bound_w_lo, bound_w_hi, db_thr, pow_thr);
        if (cls == 1) atomicMin(&s_apick[set], tid);
        else if (cls == 2) atomicMin(&s_amb[set], tid);
      }
    }
    const long long c2 = dbg ? clock64() : 0;
    __syncthreads();
    const long long c3 = dbg ? clock64() : 0;
    if (follow && s_abort) { suspended = true; break; }   // nothing of this segment has been committed

    if (has_video) {
      const long long lo = s_lo[set];
      const int pick = s_vpick[set];
      if (pick >= 0) opt = ftimes[hint + pick];      // hm:1057 (the pick is relative to this segment's hint)
      hint = lo >= 0 ? lo : nf;
    }
    if (scan_audio) {
      int apick = s_apick[set];
      const int amb = s_amb[set];
      int64_t base = first;
      // One exact step over the 64 windows below `base`, four lanes per window (every thread takes part).  The
      // windows the classifier decided come out the same way, so s_apick needs no reset.
      auto exact_step = [&]() -> int {
        __syncthreads();                              // everyone has read s_apick / s_amb
        const int64_t i = base - (int64_t)(tid / kSegLanes) * w;
        int64_t ws = 0, we = 0;
        if (i > 0) {
          ws = s0 + i; we = ws + w;
          if (ws > S.ns) ws = S.ns;
          if (we > S.ns) we = S.ns;
        }
        const double ss = group_window_sumsq(S.pcm, S.pcm_dtype, S.nch, S.e16, S.e512, ws, we, tid & (kSegLanes - 1));
        if (i > 0 && (tid & (kSegLanes - 1)) == 0 && below_level(ss, we - ws, db_thr, pow_thr)) atomicMin(&s_apick[set], tid / kSegLanes);
        __syncthreads();
        ++n_exact;
        return s_apick[set];
      };
      if (amb < apick) apick = exact_step();          // the earliest candidate is undecided
      // further steps only when the span holds more than 64 windows (max_segment_duration > 32 s)
      while (apick == 0x7fffffff && base - (int64_t)kSegWindows * w > 0) {
        base -= (int64_t)kSegWindows * w;
        apick = exact_step();
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
    if (dbg) { const long long c4 = clock64(); t_pre += c2 - c0; t_bar += c3 - c2; t_tail += c4 - c3; }
  }
  if (tid == 0) {
    *S.out_count = overflow ? -1 : count;
    if (states != nullptr) {
      states[si].current_start = cs; states[si].hint = hint; states[si].count = count;
      states[si].done = (suspended && !overflow) ? 0 : 1;
    }
  }
  if (dbg && tid == 0 && si == 0) { dbg[0] = t_pre; dbg[1] = t_bar; dbg[2] = t_tail; dbg[3] = n_exact; dbg[6] = count; }
}

}  // namespace hippo

static hippo_status launch_segment(const hippo_stream_desc* streams, int32_t nstreams, double max_segment_duration,
                                   double min_segment_duration, double frame_similarity_threshold,
                                   double audio_silence_threshold, hippo_segment_state* states, int64_t frames_ready,
                                   int final_pass, void* stream, size_t smem_reserve = 0, long long follow_ns = 0,
                                   const unsigned int* follow_gate = nullptr) {
  using namespace hippo;
  HIPPO_REQUIRE(nstreams >= 0, "hippo_segment_boundaries: nstreams < 0");
  if (nstreams == 0) return HIPPO_OK;
  HIPPO_REQUIRE(streams != nullptr, "hippo_segment_boundaries: null stream table");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  size_t smem = (size_t)2 * kStageFrames * sizeof(double);
  if (smem_reserve > smem) smem = smem_reserve;     // pattern.cu: ask for a whole SM (the extra bytes are never touched)
  static bool attr = false;
  if (!attr) { HIPPO_CUDA(cudaFuncSetAttribute(segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
  unsigned long long* dbg = nullptr;
  if (getenv("HIPPO_SEG_DEBUG")) { cudaMalloc(&dbg, 64); cudaMemset(dbg, 0, 64); }
  segment_kernel<<<nstreams, kSegThreads, smem, (cudaStream_t)stream>>>(
      streams, nstreams, max_segment_duration, min_segment_duration, frame_similarity_threshold,
      audio_silence_threshold, states, frames_ready, final_pass, follow_ns, follow_gate, dbg);
  if (dbg) {
    unsigned long long h[8];
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaMemcpy(h, dbg, 64, cudaMemcpyDeviceToHost);
    cudaFree(dbg);
    fprintf(stderr, "[seg] %llu segments, %llu exact audio steps; cycles/segment (thread 0): scan %.0f barrier %.0f tail %.0f\n", h[6], h[3],
            (double)h[0] / h[6], (double)h[1] / h[6], (double)h[2] / h[6]);
  }
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

namespace hippo {
hippo_status segment_resume_launch(const hippo_stream_desc* streams, int32_t nstreams, hippo_segment_state* states,
                                   int64_t frames_ready, int final_pass, double max_dur, double min_dur, double ssim_thr,
                                   double db_thr, size_t smem_reserve, long long follow_ns,
                                   const unsigned int* follow_gate, cudaStream_t s) {
  return launch_segment(streams, nstreams, max_dur, min_dur, ssim_thr, db_thr, states, frames_ready, final_pass, (void*)s,
                        smem_reserve, follow_ns, follow_gate);
}
int segment_stage_frames() { return kStageFrames; }
}  // namespace hippo

extern "C" hippo_status hippo_segment_boundaries(const hippo_stream_desc* streams, int32_t nstreams,
                                                 double max_segment_duration, double min_segment_duration,
                                                 double frame_similarity_threshold,
                                                 double audio_silence_threshold, void* stream) {
  return launch_segment(streams, nstreams, max_segment_duration, min_segment_duration, frame_similarity_threshold,
                        audio_silence_threshold, nullptr, 0, 1, stream);
}

extern "C" hippo_status hippo_segment_boundaries_resume(const hippo_stream_desc* streams, int32_t nstreams,
                                                        hippo_segment_state* states, int64_t frames_ready,
                                                        int32_t final_pass, double max_segment_duration,
                                                        double min_segment_duration,
                                                        double frame_similarity_threshold,
                                                        double audio_silence_threshold, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(states != nullptr, "hippo_segment_boundaries_resume: null state table");
  HIPPO_REQUIRE(frames_ready >= 0, "hippo_segment_boundaries_resume: frames_ready < 0");
  return launch_segment(streams, nstreams, max_segment_duration, min_segment_duration, frame_similarity_threshold,
                        audio_silence_threshold, states, frames_ready, final_pass ? 1 : 0, stream);
}
