// Detailed recall over the concatenated bank of ALL ThetaEvents.
//
// Reference: `_find_relevant_video_segments` / `_find_relevant_audio_segments` loop over the events of
// the long-term store (hm:3143, hm:3295) and call top_k_cosine_similarity(query, event_features, k=5)
// once per event (hm:3153, hm:3304); then every (similarity, frame) pair whose index is inside the
// event's time table becomes a +-1 s window (hm:3258-3272, hm:3366-3377), all windows are sorted by
// similarity (stable, descending) and the best five are returned (hm:3274-3277, hm:3379-3381).
//
// Here the events' feature rows live in ONE device bank (bf16 rows + fp32 norms) with an offset table:
//   hippo_scores_single    one streaming pass: score[row] = dot / (|b| * |a|) for every row of the bank
//                          (HBM bound: n * d * 2 bytes read, n * 4 written);
//   hippo_topk_segmented   the pass above + one warp per event selecting that event's k best rows;
//   hippo_recall_windows   validity test, stable global sort and window arithmetic of the loop's tail.
#include "common.cuh"

namespace hippo {

constexpr int kScoreThreads = 512;
constexpr int kScoreWarps = kScoreThreads / 32;
constexpr int kScoreRows = 4;

template <int CH>  // CH = d/256 when d is a multiple of 256 and <= 1024; 0 = generic
__global__ void __launch_bounds__(kScoreThreads, 2)
scores_single_kernel(const __nv_bfloat16* __restrict__ bank, const float* __restrict__ norm, int64_t n, int d,
                     const float* __restrict__ q, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // query permuted so that a lane's two float4 loads are conflict free (same layout as topk_single.cu):
  // element e = c*256 + lane*8 + h*4 + j  ->  sq[((c*2 + h)*32 + lane)*4 + j]
  float* sq = reinterpret_cast<float*>(smem_raw);
  const int dpad = (d + 255) / 256 * 256;
  __shared__ float s_an;
  __shared__ double s_part[kScoreWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  double ss = 0.0;
  for (int e = threadIdx.x; e < dpad; e += blockDim.x) {
    const float v = e < d ? q[e] : 0.f;
    ss += (double)v * (double)v;
    const int c = e >> 8, l = (e >> 3) & 31, h = (e >> 2) & 1, j = e & 3;
    sq[(((c * 2 + h) * 32) + l) * 4 + j] = v;
  }
  ss = warp_sum(ss);
  if (lane == 0) s_part[wid] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kScoreWarps; ++i) t += s_part[i];
    s_an = __fsqrt_rn((float)t);           // |a| as the reference forms it: fp32 sqrt of the sum of squares
  }
  __syncthreads();
  const float an = s_an;

  const int chunks = dpad >> 8;
  const int64_t ngroups = (n + kScoreRows - 1) / kScoreRows;
  const int64_t gwarp = (int64_t)blockIdx.x * kScoreWarps + wid;
  const int64_t gstride = (int64_t)gridDim.x * kScoreWarps;
  for (int64_t g = gwarp; g < ngroups; g += gstride) {
    const int64_t r0 = g * kScoreRows;
    float acc[kScoreRows];
    const __nv_bfloat16* rp[kScoreRows];
#pragma unroll
    for (int r = 0; r < kScoreRows; ++r) {
      acc[r] = 0.f;
      const int64_t row = r0 + r < n ? r0 + r : n - 1;
      rp[r] = bank + row * (int64_t)d + lane * 8;
    }
    auto fma8 = [&](const uint4 x, const float4 qa, const float4 qb, float a) {
      a = fmaf(bf16lo(x.x), qa.x, a); a = fmaf(bf16hi(x.x), qa.y, a);
      a = fmaf(bf16lo(x.y), qa.z, a); a = fmaf(bf16hi(x.y), qa.w, a);
      a = fmaf(bf16lo(x.z), qb.x, a); a = fmaf(bf16hi(x.z), qb.y, a);
      a = fmaf(bf16lo(x.w), qb.z, a); a = fmaf(bf16hi(x.w), qb.w, a);
      return a;
    };
    if constexpr (CH > 0) {
      uint4 w[CH][kScoreRows];
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int r = 0; r < kScoreRows; ++r) w[c][r] = ldg_stream(rp[r] + c * 256);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float4 qa = *reinterpret_cast<const float4*>(&sq[((c * 2 + 0) * 32 + lane) * 4]);
        const float4 qb = *reinterpret_cast<const float4*>(&sq[((c * 2 + 1) * 32 + lane) * 4]);
#pragma unroll
        for (int r = 0; r < kScoreRows; ++r) acc[r] = fma8(w[c][r], qa, qb, acc[r]);
      }
    } else {
      for (int c = 0; c < chunks; ++c) {
        if (c * 256 + lane * 8 < d) {
          const float4 qa = *reinterpret_cast<const float4*>(&sq[((c * 2 + 0) * 32 + lane) * 4]);
          const float4 qb = *reinterpret_cast<const float4*>(&sq[((c * 2 + 1) * 32 + lane) * 4]);
          uint4 w[kScoreRows];
#pragma unroll
          for (int r = 0; r < kScoreRows; ++r) w[r] = ldg_stream(rp[r] + c * 256);
#pragma unroll
          for (int r = 0; r < kScoreRows; ++r) acc[r] = fma8(w[r], qa, qb, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kScoreRows; ++r) acc[r] = warp_sum(acc[r]);
    float dot = acc[0];
#pragma unroll
    for (int r = 1; r < kScoreRows; ++r) dot = lane == r ? acc[r] : dot;
    const int64_t row = r0 + lane;
    if (lane < kScoreRows && row < n) {
      // the reference's operation order (vo:182): dot / (|b| * |a|), IEEE fp32
      out[row] = __fdiv_rn(dot, __fmul_rn(norm[row], an));
    }
  }
}

// One warp per event: k selection passes over the event's scores (order key = score desc, NaN first,
// lower row first on ties), event-LOCAL row numbers out, -1 / 0 where the event has fewer than k rows.
__global__ void __launch_bounds__(256)
segmented_select_kernel(const float* __restrict__ scores, const int64_t* __restrict__ offsets, int nseg, int k,
                        int64_t* __restrict__ out_idx, float* __restrict__ out_score,
                        float* __restrict__ out_max) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= nseg) return;
  const int64_t lo = offsets[e], hi = offsets[e + 1];
  uint64_t prev = ~0ull;
  for (int r = 0; r < k; ++r) {
    uint64_t best = 0;
    if (prev != 0) {
      for (int64_t i = lo + lane; i < hi; i += 32) {
        const uint64_t c = pack_key(scores[i], (uint32_t)(i - lo));
        if (c < prev && c > best) best = c;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
    }
    if (lane == 0) {
      const size_t o = (size_t)e * k + r;
      out_idx[o] = best ? (int64_t)key_row(best) : -1;
      out_score[o] = best ? key_score(best) : 0.f;
      // np.max over the event's top-k (hm:3156, hm:3307): the first entry, NaN if any score is NaN
      if (r == 0 && out_max) out_max[e] = best ? key_score(best) : __int_as_float(0x7fc00000);
    }
    prev = best;
  }
}

// Tail of the recall loop.  Candidate c = e * k + r is live iff its event is enabled, idx >= 0 and
// idx < (number of times of event e) (hm:3262, hm:3367); the m best by (score desc, c asc) -- Python's
// stable sort with reverse=True (hm:3274, hm:3379) -- are written with their windows
// [max(0, t - pad), t + pad] (hm:3265-3266).  One CTA; candidates are few (events x 5).
__global__ void __launch_bounds__(256)
recall_windows_kernel(const int64_t* __restrict__ seg_idx, const float* __restrict__ seg_score, int nseg, int k,
                      const int64_t* __restrict__ time_offsets, const double* __restrict__ times,
                      const uint8_t* __restrict__ enabled, double pad, int m, int32_t* __restrict__ out_event,
                      int64_t* __restrict__ out_idx, float* __restrict__ out_score,
                      double* __restrict__ out_window, int32_t* __restrict__ out_count) {
  __shared__ uint64_t s_best[8];
  __shared__ uint64_t s_prev;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t ncand = (int64_t)nseg * k;
  if (threadIdx.x == 0) s_prev = ~0ull;
  __syncthreads();
  int count = 0;
  for (int r = 0; r < m; ++r) {
    const uint64_t prev = s_prev;
    uint64_t best = 0;
    if (prev != 0) {
      for (int64_t c = threadIdx.x; c < ncand; c += blockDim.x) {
        const int e = (int)(c / k);
        const int64_t idx = seg_idx[c];
        if (enabled && !enabled[e]) continue;
        if (idx < 0 || idx >= time_offsets[e + 1] - time_offsets[e]) continue;
        const uint64_t key = pack_key(seg_score[c], (uint32_t)c);   // ties: lower candidate number first
        if (key < prev && key > best) best = key;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
    }
    if (lane == 0) s_best[wid] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t b = 0;
      for (int i = 0; i < 8; ++i) b = s_best[i] > b ? s_best[i] : b;
      s_prev = b;
      if (b) {
        const int64_t c = (int64_t)key_row(b);
        const int e = (int)(c / k);
        const int64_t idx = seg_idx[c];
        const double t = times[time_offsets[e] + idx];
        out_event[r] = e;
        out_idx[r] = idx;
        out_score[r] = seg_score[c];
        const double lo = t - pad;
        out_window[2 * r] = lo > 0.0 ? lo : 0.0;     // max(0, t - 1): a NaN time gives 0 in Python too
        out_window[2 * r + 1] = t + pad;
        ++count;
      } else {
        out_event[r] = -1;
        out_idx[r] = -1;
        out_score[r] = 0.f;
        out_window[2 * r] = 0.0;
        out_window[2 * r + 1] = 0.0;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_count = count;
}

static hippo_status launch_scores(const void* bank, const float* norm, int64_t n, int d, const float* q,
                                  float* out, cudaStream_t s) {
  const size_t smem = (size_t)((d + 255) / 256 * 256) * 4;
  HIPPO_REQUIRE(smem <= 200 * 1024, "scores: d=%d too large", d);
  const int64_t groups = (n + kScoreRows - 1) / kScoreRows;
  int64_t want = (groups + kScoreWarps - 1) / kScoreWarps;
  const int64_t cap = (int64_t)sm_count() * 2;
  if (want < 1) want = 1;
  const int grid = (int)(want < cap ? want : cap);
  const __nv_bfloat16* b = (const __nv_bfloat16*)bank;
  if (d == 1024) {
    HIPPO_CUDA(cudaFuncSetAttribute(scores_single_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    scores_single_kernel<4><<<grid, kScoreThreads, smem, s>>>(b, norm, n, d, q, out);
  } else {
    HIPPO_CUDA(cudaFuncSetAttribute(scores_single_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    scores_single_kernel<0><<<grid, kScoreThreads, smem, s>>>(b, norm, n, d, q, out);
  }
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

}  // namespace hippo

extern "C" {

hippo_status hippo_scores_single(const void* bank, const float* norm, int64_t n, int32_t d, const float* q,
                                 float* out_score, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(n >= 0 && d > 0 && d % 64 == 0, "hippo_scores_single: need d %% 64 == 0 (d=%d)", d);
  if (n == 0) return HIPPO_OK;
  HIPPO_REQUIRE(bank && norm && q && out_score, "hippo_scores_single: null pointer");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  return launch_scores(bank, norm, n, d, q, out_score, (cudaStream_t)stream);
}

size_t hippo_topk_segmented_workspace_bytes(int64_t n) {
  return hippo::align_up((size_t)(n > 0 ? n : 1) * sizeof(float), 256);
}

hippo_status hippo_topk_segmented(const void* bank, const float* norm, int64_t n, int32_t d, const float* q,
                                  const int64_t* seg_offsets, int32_t nseg, int32_t k, int64_t* out_idx,
                                  float* out_score, float* out_max, void* ws, size_t ws_bytes, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(n >= 0 && d > 0 && d % 64 == 0, "hippo_topk_segmented: need d %% 64 == 0 (d=%d)", d);
  HIPPO_REQUIRE(nseg >= 0 && k >= 1, "hippo_topk_segmented: bad nseg / k");
  HIPPO_REQUIRE(n < 0xffffffffll, "hippo_topk_segmented: n must stay below 2^32-1");
  if (nseg == 0) return HIPPO_OK;
  HIPPO_REQUIRE(q && seg_offsets && out_idx && out_score && (n == 0 || (bank && norm)),
                "hippo_topk_segmented: null pointer");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  if (ws == nullptr || ws_bytes < (size_t)n * sizeof(float) || ((uintptr_t)ws & 255)) {
    set_error("hippo_topk_segmented: workspace of %zu bytes needed (256-byte aligned)",
              hippo_topk_segmented_workspace_bytes(n));
    return HIPPO_E_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  float* scores = (float*)ws;
  if (n > 0) {
    st = launch_scores(bank, norm, n, d, q, scores, s);
    if (st != HIPPO_OK) return st;
  }
  segmented_select_kernel<<<(nseg + 7) / 8, 256, 0, s>>>(scores, seg_offsets, nseg, k, out_idx, out_score, out_max);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

hippo_status hippo_recall_windows(const int64_t* seg_idx, const float* seg_score, int32_t nseg, int32_t k,
                                  const int64_t* time_offsets, const double* times, const uint8_t* enabled,
                                  double pad, int32_t m, int32_t* out_event, int64_t* out_idx, float* out_score,
                                  double* out_window, int32_t* out_count, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(nseg >= 0 && k >= 1 && m >= 1, "hippo_recall_windows: bad sizes");
  HIPPO_REQUIRE((int64_t)nseg * k < 0xffffffffll, "hippo_recall_windows: too many candidates");
  HIPPO_REQUIRE(out_event && out_idx && out_score && out_window && out_count, "hippo_recall_windows: null output");
  HIPPO_REQUIRE(nseg == 0 || (seg_idx && seg_score && time_offsets && times), "hippo_recall_windows: null input");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  recall_windows_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(seg_idx, seg_score, nseg, k, time_offsets, times,
                                                            enabled, pad, m, out_event, out_idx, out_score,
                                                            out_window, out_count);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

}  // extern "C"
