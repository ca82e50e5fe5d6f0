// hippo_pattern_separation: temporal pattern separation of ONE stream (hm:1002-1114 with its helpers hm:980-1000),
// every stage of it, issued from ONE call with the stages overlapped on the device.
//
// The stages have different bounds -- BGR -> gray is HBM bound, SSIM is integer-issue bound, the audio pyramid is HBM
// bound and tiny, the boundary state machine is one sequential latency chain -- so run back to back they leave most
// of the machine idle most of the time (0.82 ms per stream-hour, 0.12 of the HBM roofline).  Here the frames are
// taken in chunks of `chunk_pairs` adjacent pairs:
//   side stream A / B (alternating)   gray + SSIM of chunk i: the gray conversion of chunk i + 1 runs under the SSIM
//                                     kernel of chunk i;
//   side stream C                     audio pyramid first, then the boundary state machine in its RESUMABLE form
//                                     (segment.cu): after every second chunk it takes the segments whose window is
//                                     covered by finished SSIM values and suspends, so the chain runs under the SSIM
//                                     kernels of the later chunks and only its last few segments remain at the end.
// Ordering is by stream events only -- no kernel ever waits for another kernel -- and the launches come from this
// host function (a Python loop over the chunks costs more host time than the kernels take).
#include "common.cuh"

#include <cstdlib>
#include <vector>

namespace hippo {

// stream descriptor + fresh chain state; clears the item counters of the SSIM launches; follow mode (pair_done != null):
// clears the per-pair partial counts and marks every pair's SSIM as pending
__global__ void pattern_init_kernel(hippo_stream_desc* desc, hippo_segment_state* state, hippo_stream_desc value,
                                    unsigned int* counters, int ncounters, unsigned int* pair_done, double* ssim,
                                    int npairs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (t == 0) {
    *desc = value;
    state->current_start = 0.0;
    state->hint = 0;
    state->count = 0;
    state->done = 0;
  }
  for (int i = t; i < ncounters; i += nt) counters[i] = 0;
  if (pair_done != nullptr)
    for (int i = t; i < npairs; i += nt) { pair_done[i] = 0; ssim[i] = __longlong_as_double((long long)kSsimPending); }
}

hippo_status frames_gray_launch(const uint8_t* frames, int nf, int h, int w, int ch, void* ws, size_t ws_bytes,
                                cudaStream_t s);
hippo_status frames_ssim_launch(int nf, int h, int w, void* ws, size_t ws_bytes, int p0, int p1, int bh,
                                unsigned int* counter, unsigned int* pair_done, double* out_ssim, double* out_mse,
                                cudaStream_t s);
hippo_status segment_resume_launch(const hippo_stream_desc* streams, int32_t nstreams, hippo_segment_state* states,
                                   int64_t frames_ready, int final_pass, double max_dur, double min_dur, double ssim_thr,
                                   double db_thr, size_t smem_reserve, long long follow_ns, cudaStream_t s);
int segment_stage_frames();
constexpr int kPatternMaxChunks = 256;

struct PatternLayout {
  void* lane_ws[2]; size_t lane_bytes;
  hippo_stream_desc* desc; hippo_segment_state* state;
  unsigned int* counters;
  unsigned int* pair_done;
  size_t bytes;
};

static bool pattern_follow(int nf) {
  const char* e = getenv("HIPPO_PATTERN_FOLLOW");     // 0: the chunk-by-chunk resumable chain of round 2's first version
  return nf > 1 && nf <= segment_stage_frames() && !(e && atoi(e) == 0);
}

static int pattern_chunk(int nf, int chunk_pairs) {
  // one wave of SSIM CTAs at 224 x 224 (8 items per pair, 8 warps per CTA): 444 pairs = three CTAs per SM on 148 SMs,
  // 588 = four per SM on the 147 SMs the follow-mode chain leaves to them
  int cp = chunk_pairs > 0 ? chunk_pairs : (pattern_follow(nf) ? 588 : 444);
  if (cp > nf - 1) cp = nf - 1;
  if (cp < 1) cp = 1;
  if ((nf - 1 + cp - 1) / cp > kPatternMaxChunks - 2) cp = (nf - 1 + kPatternMaxChunks - 3) / (kPatternMaxChunks - 2);
  return cp;
}

static PatternLayout pattern_layout(void* ws, size_t ws_bytes, int nf, int h, int w, int chunk_pairs) {
  Carver c(ws, ws_bytes);
  PatternLayout L{};
  const int cp = pattern_chunk(nf, chunk_pairs);
  (void)cp;
  // one frame-pair workspace for the whole stream: gray image of every frame, per-pair partial sums
  L.lane_bytes = nf > 1 ? align_up(hippo_frame_pairs_workspace_bytes(nf, h, w, nf - 1), 256) : 256;
  L.lane_ws[0] = c.take<unsigned char>(L.lane_bytes);
  L.lane_ws[1] = nullptr;
  L.desc = c.take<hippo_stream_desc>(1);
  L.state = c.take<hippo_segment_state>(1);
  L.counters = c.take<unsigned int>(kPatternMaxChunks);
  L.pair_done = c.take<unsigned int>(nf > 1 ? nf - 1 : 1);
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
  if (nf < 0 || h <= 0 || w <= 0) return 256;
  return hippo::pattern_layout(nullptr, 0, nf, h, w, chunk_pairs).bytes;
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
  const bool has_video = frames != nullptr && frame_times != nullptr && nf > 0;
  const bool has_audio = pcm != nullptr;
  HIPPO_REQUIRE(nf >= 0 && (!has_video || (h >= 1 && w >= 1 && (ch == 1 || ch == 3))), "hippo_pattern_separation: bad frame shape");
  HIPPO_REQUIRE(!has_video || nf < 2 || (out_ssim && out_mse), "hippo_pattern_separation: out_ssim / out_mse missing");
  HIPPO_REQUIRE(!has_audio || (ns >= 0 && nch >= 1 && out_e16 && out_e512), "hippo_pattern_separation: bad audio arguments");
  HIPPO_REQUIRE(out_bounds && out_count && max_segments >= 1, "hippo_pattern_separation: bad output arguments");
  HIPPO_REQUIRE(side_streams_host != nullptr, "hippo_pattern_separation: three side streams are needed");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  const int nfv = has_video ? nf : 0;
  PatternLayout L = pattern_layout(ws, ws_bytes, nfv, h, w, chunk_pairs);
  if (ws == nullptr || ((uintptr_t)ws & 255) || L.bytes > ws_bytes) {
    set_error("hippo_pattern_separation: workspace of %zu bytes needed (256-byte aligned), got %zu", L.bytes, ws_bytes);
    return HIPPO_E_WORKSPACE;
  }
  cudaStream_t main = (cudaStream_t)stream;
  cudaStream_t lane[2] = {(cudaStream_t)side_streams_host[0], (cudaStream_t)side_streams_host[1]};
  cudaStream_t chain = (cudaStream_t)side_streams_host[2];
  const int cp = pattern_chunk(nfv, chunk_pairs);
  // SSIM chunks: whole chunks of cp pairs (444 = one wave of SSIM CTAs); a small remainder stays its own last chunk:
  // it runs beside the chunk before it (alternating streams), and the shorter the last chunk, the fewer segments are
  // left for the final pass of the chain -- the only part of the chain nothing can hide
  std::vector<int> cuts;                       // first pair of every chunk, then the number of pairs
  if (nfv > 1) {
    for (int f = 0; f < nfv - 1; f += cp) cuts.push_back(f);
    cuts.push_back(nfv - 1);
  }
  const int nchunks = cuts.empty() ? 0 : (int)cuts.size() - 1;
  size_t ev = 0;
  cudaEvent_t e_start = pooled_event(ev++);
  HIPPO_REQUIRE(e_start != nullptr, "hippo_pattern_separation: could not create events");
  HIPPO_CUDA(cudaEventRecord(e_start, main));
  HIPPO_CUDA(cudaStreamWaitEvent(lane[0], e_start, 0));
  HIPPO_CUDA(cudaStreamWaitEvent(chain, e_start, 0));

  // HIPPO_PATTERN_DEBUG: a timeline of the chunks and of the chain launches (timing events, synchronises at the end)
  const bool dbg = getenv("HIPPO_PATTERN_DEBUG") != nullptr;
  struct Mark { cudaEvent_t e; const char* what; int i; };
  std::vector<Mark> marks;
  auto mark = [&](cudaStream_t s, const char* what, int i) {
    if (!dbg) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    marks.push_back({e, what, i});
  };
  mark(main, "start", 0);

  // chain stream: audio pyramid, stream descriptor + fresh state
  if (has_audio) {
    st = hippo_audio_energy(pcm, pcm_dtype, ns, nch, out_e16, out_e512, chain);
    if (st != HIPPO_OK) return st;
  }
  hippo_stream_desc d{};
  d.ssim = nfv > 1 ? out_ssim : nullptr;
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
  const bool follow = nchunks > 0 && pattern_follow(nfv);
  pattern_init_kernel<<<follow ? 8 : 1, 256, 0, chain>>>(L.desc, L.state, d, L.counters, kPatternMaxChunks,
                                                         follow ? L.pair_done : nullptr, out_ssim, nfv - 1);
  HIPPO_CUDA(cudaGetLastError());
  cudaEvent_t e_init = pooled_event(ev++);
  HIPPO_REQUIRE(e_init != nullptr, "hippo_pattern_separation: could not create events");
  HIPPO_CUDA(cudaEventRecord(e_init, chain));
  const size_t chain_smem = 0;
  if (follow) {
    // the chain, ONE launch that follows the SSIM kernels pair by pair from an SM of its own (segment.cu): 220 KB of
    // dynamic shared memory leave no room for a gray (12 KB static) or SSIM (8 KB dynamic) CTA beside it.  A wait
    // beyond the limit (HIPPO_FOLLOW_US, default 4 ms) suspends it; the final pass below picks up whatever is left.
    const char* fe = getenv("HIPPO_FOLLOW_US");
    const long long follow_ns = 1000ll * (fe && atoll(fe) > 0 ? atoll(fe) : 4000);
    mark(chain, "follow begin", 0);
    st = segment_resume_launch(L.desc, 1, L.state, 0, 0, max_segment_duration, min_segment_duration,
                               frame_similarity_threshold, audio_silence_threshold, (size_t)220 * 1024, follow_ns, chain);
    if (st != HIPPO_OK) return st;
    mark(chain, "follow end", 0);
  }
  // persistent SSIM warps (dynamic item queue) finish the SSIM phase sooner (444 vs 483 us per stream-hour) but crowd
  // the boundary chain, which then ends later: off unless HIPPO_SSIM_PERSISTENT=1
  const bool persistent = getenv("HIPPO_SSIM_PERSISTENT") && atoi(getenv("HIPPO_SSIM_PERSISTENT")) == 1;
  const char* fine_env = getenv("HIPPO_SSIM_FINE");
  const int fine_bh = fine_env ? atoi(fine_env) : 56;

  if (nchunks > 0) {
    // lane 0: BGR -> gray (+ min / max) of ALL frames, once.  It is HBM bound; the SSIM kernels are issue bound, but
    // letting them overlap does not pay: the conversion streams 0.7 GB through L2 and evicts the gray rows the SSIM
    // warps are reading (measured: chunks of gray + SSIM on alternating streams were slower than no overlap at all)
    st = frames_gray_launch(frames, nfv, h, w, ch, L.lane_ws[0], L.lane_bytes, lane[0]);
    if (st != HIPPO_OK) return st;
    cudaEvent_t e_gray = pooled_event(ev++);
    HIPPO_REQUIRE(e_gray != nullptr, "hippo_pattern_separation: could not create events");
    HIPPO_CUDA(cudaEventRecord(e_gray, lane[0]));
    HIPPO_CUDA(cudaStreamWaitEvent(lane[1], e_gray, 0));
    HIPPO_CUDA(cudaStreamWaitEvent(lane[0], e_init, 0));     // counters cleared
    HIPPO_CUDA(cudaStreamWaitEvent(lane[1], e_init, 0));
    mark(lane[0], "gray end", 0);
  }
  for (int i = 0; i < nchunks; ++i) {
    // SSIM of pairs [p0, p1) on alternating streams: the CTAs of chunk i + 1 fill the SMs as chunk i drains
    const int p0 = cuts[i], p1 = cuts[i + 1];
    cudaStream_t s = lane[i & 1];
    // the chunks that end the stream use bands half as high: twice as many, shorter items, so the SMs run
    // dry within ~25 us of each other instead of ~50 (everything before is followed by more work anyway)
    const int bh = (nfv - 1 - p0 <= cp + cp / 2) ? fine_bh : 56;
    st = frames_ssim_launch(nfv, h, w, L.lane_ws[0], L.lane_bytes, p0, p1, bh,
                            (persistent && !follow) ? L.counters + i : nullptr, follow ? L.pair_done : nullptr, out_ssim,
                            out_mse, s);
    if (st != HIPPO_OK) return st;
    mark(s, "chunk end", i);
    const bool last = i == nchunks - 1;
    if (follow && i < nchunks - 2) continue;       // streams are in order: the last event of either lane covers the lane
    cudaEvent_t e = pooled_event(ev++);
    HIPPO_REQUIRE(e != nullptr, "hippo_pattern_separation: could not create events");
    HIPPO_CUDA(cudaEventRecord(e, s));
    HIPPO_CUDA(cudaStreamWaitEvent(chain, e, 0));
    if (follow && !last) continue;                 // one final pass behind everything (a no-op when the chain got through)
    mark(chain, "chain begin", i);
    // pairs < p1 are final, i.e. frames < p1 + 1 are covered (every earlier chunk's event was waited for above)
    st = segment_resume_launch(L.desc, 1, L.state, p1 + 1, last ? 1 : 0, max_segment_duration, min_segment_duration,
                               frame_similarity_threshold, audio_silence_threshold, chain_smem, 0, chain);
    if (st != HIPPO_OK) return st;
    mark(chain, "chain end", i);
  }
  if (nchunks == 0) {
    st = segment_resume_launch(L.desc, 1, L.state, nfv, 1, max_segment_duration, min_segment_duration,
                               frame_similarity_threshold, audio_silence_threshold, 0, 0, chain);
    if (st != HIPPO_OK) return st;
  }
  cudaEvent_t e_done = pooled_event(ev++);
  HIPPO_REQUIRE(e_done != nullptr, "hippo_pattern_separation: could not create events");
  HIPPO_CUDA(cudaEventRecord(e_done, chain));
  HIPPO_CUDA(cudaStreamWaitEvent(main, e_done, 0));
  if (dbg) {
    mark(main, "joined", 0);
    cudaStreamSynchronize(main);
    for (size_t i = 1; i < marks.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[0].e, marks[i].e);
      fprintf(stderr, "[pattern] %8.1f us  %s %d\n", ms * 1e3f, marks[i].what, marks[i].i);
    }
    for (auto& m : marks) cudaEventDestroy(m.e);
  }
  return HIPPO_OK;
}

}  // extern "C"
