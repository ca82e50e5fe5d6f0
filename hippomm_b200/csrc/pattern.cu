// hippo_pattern_separation: temporal pattern separation of ONE stream (hm:1002-1114 with its helpers hm:980-1000),
// every stage of it, issued from ONE call with the stages overlapped on the device.
//
// The stages have different bounds -- BGR -> gray is HBM bound, SSIM is integer-issue bound, the audio pyramid is HBM
// bound and tiny, the boundary state machine is one sequential latency chain -- so run back to back they leave most
// of the machine idle most of the time (0.82 ms per stream-hour, 0.12 of the HBM roofline).  Here:
//   side stream 2   the boundary state machine in FOLLOW mode (segment.cu), launched FIRST: one CTA on an SM of its
//                   own that polls the SSIM values it needs as they are delivered, so the chain ends a few
//                   microseconds after the last pair; then a final resumable pass behind everything, which completes
//                   the chain if the follower gave up (it never waits longer than its limit, so a serialising
//                   profiler or a launch-blocking debug run only costs that limit);
//   side stream 1   audio pyramid (+ a flag the follower's gate polls);
//   side stream 0   gray conversion, then the SSIM of all adjacent pairs as one launch, every pair's result stored by
//                   the warp that completes it (frames.cu, live finalisation).
// The launches come from this host function (a Python loop costs more host time than the kernels take).
#include "common.cuh"

#include <cstdlib>
#include <vector>

namespace hippo {

// stream descriptor + fresh chain state; follow mode (pair_done != null): clears the per-pair partial counts and
// marks every pair's SSIM as pending
__global__ void pattern_init_kernel(hippo_stream_desc* desc, hippo_segment_state* state, hippo_stream_desc value,
                                    unsigned int* pair_done, double* ssim, int npairs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (t == 0) {
    *desc = value;
    state->current_start = 0.0;
    state->hint = 0;
    state->count = 0;
    state->done = 0;
  }
  if (pair_done != nullptr) {
    for (int i = t; i < npairs; i += nt) { pair_done[i] = 0; ssim[i] = __longlong_as_double((long long)kSsimPending); }
    if (t == 0) { pair_done[npairs] = 0; pair_done[npairs + 1] = 0; }
  }
}

__global__ void pattern_flag_kernel(unsigned int* flag) { __threadfence(); *flag = 1u; }
// holds a stream back until the follower is resident (or `limit_ns` have passed: it may be running AFTER us under a
// serialising tool)
__global__ void pattern_wait_kernel(const unsigned int* flag, long long limit_ns) {
  const unsigned long long t0 = globaltimer_ns();
  while (ld_volatile_u32(flag) == 0 && globaltimer_ns() - t0 < (unsigned long long)limit_ns) __nanosleep(100);
}

hippo_status frames_adjacent_live_launch(const uint8_t* frames, int nf, int h, int w, int ch, void* ws, size_t ws_bytes,
                                         unsigned int* pair_done, double* out_ssim, double* out_mse, cudaStream_t s);
hippo_status segment_resume_launch(const hippo_stream_desc* streams, int32_t nstreams, hippo_segment_state* states,
                                   int64_t frames_ready, int final_pass, double max_dur, double min_dur, double ssim_thr,
                                   double db_thr, size_t smem_reserve, long long follow_ns,
                                   const unsigned int* follow_gate, cudaStream_t s);
int segment_stage_frames();
struct PatternLayout {
  void* frame_ws; size_t frame_bytes;
  hippo_stream_desc* desc; hippo_segment_state* state;
  unsigned int* pair_done;                    // [npairs] partials delivered per pair, then "audio pyramid complete", "follower resident"
  size_t bytes;
};

static bool pattern_follow() {
  const char* e = getenv("HIPPO_PATTERN_FOLLOW");     // 0: no follower, the final pass runs the whole chain (A/B, debugging)
  return !(e && atoi(e) == 0);
}

static PatternLayout pattern_layout(void* ws, size_t ws_bytes, int nf, int h, int w) {
  Carver c(ws, ws_bytes);
  PatternLayout L{};
  // one frame-pair workspace for the whole stream: gray image of every frame, per-pair partial sums
  L.frame_bytes = nf > 1 ? align_up(hippo_frame_pairs_workspace_bytes(nf, h, w, nf - 1), 256) : 256;
  L.frame_ws = c.take<unsigned char>(L.frame_bytes);
  L.desc = c.take<hippo_stream_desc>(1);
  L.state = c.take<hippo_segment_state>(1);
  L.pair_done = c.take<unsigned int>((nf > 1 ? nf - 1 : 0) + 2);
  L.bytes = c.used();
  return L;
}

// events are created once per host thread and reused: recording an event again only affects later waits
static cudaEvent_t pooled_event(size_t i) {
  static thread_local std::vector<cudaEvent_t> pool;
  while (pool.size() <= i) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    pool.push_back(e);
  }
  return pool[i];
}

}  // namespace hippo

extern "C" {

size_t hippo_pattern_separation_workspace_bytes(int32_t nf, int32_t h, int32_t w, int32_t chunk_pairs) {
  (void)chunk_pairs;
  if (nf < 0 || h <= 0 || w <= 0) return 256;
  return hippo::pattern_layout(nullptr, 0, nf, h, w).bytes;
}

hippo_status hippo_pattern_separation(const uint8_t* frames, int32_t nf, int32_t h, int32_t w, int32_t ch,
                                      const double* frame_times, const void* pcm, int32_t pcm_dtype, int64_t ns,
                                      int32_t nch, double sample_rate, double max_segment_duration,
                                      double min_segment_duration, double frame_similarity_threshold,
                                      double audio_silence_threshold, int32_t chunk_pairs, double* out_ssim,
                                      double* out_mse, double* out_e16, double* out_e512, double* out_bounds,
                                      int32_t* out_count, int32_t max_segments, void* ws, size_t ws_bytes,
                                      void* const* side_streams_host, void* stream) {
  using namespace hippo;
  (void)chunk_pairs;                              // the first version took the frames in chunks; one SSIM launch now
  const bool has_video = frames != nullptr && frame_times != nullptr && nf > 0;
  const bool has_audio = pcm != nullptr;
  HIPPO_REQUIRE(nf >= 0 && (!has_video || (h >= 1 && w >= 1 && (ch == 1 || ch == 3))), "hippo_pattern_separation: bad frame shape");
  HIPPO_REQUIRE(!has_video || (nf <= 65535 && w <= 65535 && h <= 65535), "hippo_pattern_separation: at most 65535 frames of 65535 x 65535");
  HIPPO_REQUIRE(!has_video || nf < 2 || (out_ssim && out_mse), "hippo_pattern_separation: out_ssim / out_mse missing");
  HIPPO_REQUIRE(!has_audio || (ns >= 0 && nch >= 1 && out_e16 && out_e512), "hippo_pattern_separation: bad audio arguments");
  HIPPO_REQUIRE(out_bounds && out_count && max_segments >= 1, "hippo_pattern_separation: bad output arguments");
  HIPPO_REQUIRE(side_streams_host != nullptr, "hippo_pattern_separation: three side streams are needed");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  const int nfv = has_video ? nf : 0;
  PatternLayout L = pattern_layout(ws, ws_bytes, nfv, h, w);
  if (ws == nullptr || ((uintptr_t)ws & 255) || L.bytes > ws_bytes) {
    set_error("hippo_pattern_separation: workspace of %zu bytes needed (256-byte aligned), got %zu", L.bytes, ws_bytes);
    return HIPPO_E_WORKSPACE;
  }
  cudaStream_t main = (cudaStream_t)stream;
  cudaStream_t lane = (cudaStream_t)side_streams_host[0];
  cudaStream_t aux = (cudaStream_t)side_streams_host[1];
  cudaStream_t chain = (cudaStream_t)side_streams_host[2];
  const bool scored = nfv > 1;                                  // there are adjacent pairs to score
  const bool follow = scored && pattern_follow();
  size_t ev = 0;
  cudaEvent_t e_start = pooled_event(ev++), e_init = pooled_event(ev++), e_pairs = pooled_event(ev++),
              e_audio = pooled_event(ev++), e_done = pooled_event(ev++);
  HIPPO_REQUIRE(e_start && e_init && e_pairs && e_audio && e_done, "hippo_pattern_separation: could not create events");
  HIPPO_CUDA(cudaEventRecord(e_start, main));
  HIPPO_CUDA(cudaStreamWaitEvent(chain, e_start, 0));

  // HIPPO_PATTERN_DEBUG: a timeline of the launches (timing events, synchronises at the end)
  const bool dbg = getenv("HIPPO_PATTERN_DEBUG") != nullptr;
  struct Mark { cudaEvent_t e; const char* what; };
  std::vector<Mark> marks;
  auto mark = [&](cudaStream_t s, const char* what) {
    if (!dbg) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    marks.push_back({e, what});
  };
  mark(main, "start");

  // The chain reads the 512-sample sums (0.9 MB per stream-hour) while the frame kernels stream ~0.9 GB through L2:
  // left alone every such read goes to DRAM behind that traffic and a segment takes ~3 us instead of 1.4.  The pyramid
  // kernels (aux) and the chain (chain) therefore access the sums with the persisting L2 policy.
  const char* pin_env = getenv("HIPPO_PATTERN_L2PIN");      // 0 switches it off (A/B)
  bool l2_window = false;
  if (follow && has_audio && !(pin_env && atoi(pin_env) == 0)) {
    static thread_local int l2_ready = 0;            // 0 = not tried, 1 = set-aside configured, -1 = unavailable
    if (l2_ready == 0) {
      int dev = 0, max_persist = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
      size_t want = (size_t)8 << 20;
      if ((size_t)max_persist < want) want = (size_t)max_persist;
      l2_ready = (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) ? 1 : -1;
      cudaGetLastError();
    }
    if (l2_ready == 1) {
      int dev = 0, max_win = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
      size_t bytes = (size_t)((ns + 511) / 512) * sizeof(double);
      if (bytes > (size_t)max_win) bytes = (size_t)max_win;
      cudaStreamAttrValue av{};
      av.accessPolicyWindow.base_ptr = out_e512;
      av.accessPolicyWindow.num_bytes = bytes;
      av.accessPolicyWindow.hitRatio = 1.0f;
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(aux, cudaStreamAttributeAccessPolicyWindow, &av);
      cudaStreamSetAttribute(chain, cudaStreamAttributeAccessPolicyWindow, &av);
      cudaGetLastError();
      l2_window = true;
    }
  }

  // chain stream: stream descriptor + fresh state (+ pending marks), then the follower at once
  hippo_stream_desc d{};
  d.ssim = scored ? out_ssim : nullptr;
  d.frame_times = has_video ? frame_times : nullptr;
  d.nframes = nfv;
  d.pcm = pcm;
  d.e16 = has_audio ? out_e16 : nullptr;
  d.e512 = has_audio ? out_e512 : nullptr;
  d.ns = has_audio ? ns : 0;
  d.nch = has_audio ? nch : 1;
  d.pcm_dtype = has_audio ? pcm_dtype : HIPPO_F64;
  d.sample_rate = has_audio ? sample_rate : 0.0;
  d.out_bounds = out_bounds;
  d.out_count = out_count;
  d.max_segments = max_segments;
  unsigned int* audio_flag = L.pair_done + (scored ? nfv - 1 : 0);
  pattern_init_kernel<<<scored ? 8 : 1, 256, 0, chain>>>(L.desc, L.state, d, scored ? L.pair_done : nullptr, out_ssim,
                                                         nfv - 1);
  HIPPO_CUDA(cudaGetLastError());
  HIPPO_CUDA(cudaEventRecord(e_init, chain));
  if (follow) {
    // the chain, ONE launch that follows the SSIM warps pair by pair from an SM of its own (segment.cu).  It goes
    // first: the GPU is idle, so it is resident before the frame kernels fill every SM, and 220 KB of dynamic shared
    // memory leave no room beside it for their CTAs (12 KB static / 8 KB dynamic).  A wait beyond the limit
    // (HIPPO_FOLLOW_US, default 4 ms) suspends it; the final pass below picks up whatever is left.
    const char* fe = getenv("HIPPO_FOLLOW_US");
    const long long follow_ns = 1000ll * (fe && atoll(fe) > 0 ? atoll(fe) : 4000);
    st = segment_resume_launch(L.desc, 1, L.state, 0, 0, max_segment_duration, min_segment_duration,
                               frame_similarity_threshold, audio_silence_threshold, (size_t)220 * 1024, follow_ns,
                               audio_flag, chain);
    if (st != HIPPO_OK) return st;
    mark(chain, "follow end");
  }
  if (has_audio) {
    // before the frame kernels in launch order: behind them it would wait for an SM until they drain
    HIPPO_CUDA(cudaStreamWaitEvent(aux, e_init, 0));
    st = hippo_audio_energy(pcm, pcm_dtype, ns, nch, out_e16, out_e512, aux);
    if (st != HIPPO_OK) return st;
    if (follow) pattern_flag_kernel<<<1, 1, 0, aux>>>(audio_flag);
    mark(aux, "pyramid end");
    HIPPO_CUDA(cudaEventRecord(e_audio, aux));
  }
  if (scored) {
    // lane: gray + SSIM of every adjacent pair behind the pending marks, held back until the follower is resident
    HIPPO_CUDA(cudaStreamWaitEvent(lane, e_init, 0));
    if (follow) pattern_wait_kernel<<<1, 1, 0, lane>>>(audio_flag + 1, 100000);
    st = frames_adjacent_live_launch(frames, nfv, h, w, ch, L.frame_ws, L.frame_bytes, L.pair_done, out_ssim, out_mse, lane);
    if (st != HIPPO_OK) return st;
    mark(lane, "pairs end");
    HIPPO_CUDA(cudaEventRecord(e_pairs, lane));
  }
  // final pass behind everything: a no-op when the follower got through
  if (scored) HIPPO_CUDA(cudaStreamWaitEvent(chain, e_pairs, 0));
  if (has_audio) HIPPO_CUDA(cudaStreamWaitEvent(chain, e_audio, 0));
  st = segment_resume_launch(L.desc, 1, L.state, nfv, 1, max_segment_duration, min_segment_duration,
                             frame_similarity_threshold, audio_silence_threshold, 0, 0, nullptr, chain);
  if (st != HIPPO_OK) return st;
  mark(chain, "final pass end");
  HIPPO_CUDA(cudaEventRecord(e_done, chain));
  HIPPO_CUDA(cudaStreamWaitEvent(main, e_done, 0));
  if (l2_window) {
    // the window applied to the launches above; whatever the caller runs on these streams next is not ours to mark
    cudaStreamAttrValue off{};
    cudaStreamSetAttribute(aux, cudaStreamAttributeAccessPolicyWindow, &off);
    cudaStreamSetAttribute(chain, cudaStreamAttributeAccessPolicyWindow, &off);
    cudaGetLastError();
  }
  if (dbg) {
    mark(main, "joined");
    cudaStreamSynchronize(main);
    for (size_t i = 1; i < marks.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[0].e, marks[i].e);
      fprintf(stderr, "[pattern] %8.1f us  %s\n", ms * 1e3f, marks[i].what);
    }
    for (auto& m : marks) cudaEventDestroy(m.e);
  }
  return HIPPO_OK;
}

}  // extern "C"
