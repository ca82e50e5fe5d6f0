// hippo_consolidate: greedy cosine-redundancy filter (reference: _select_key_frames, hm:944-967).
//
//   1. hippo_bank_build      fp32 rows -> bf16 rows + fp32 norms              (hm:951)
//   2. sim_tc EPI_MASK       lower-triangle similarity bits on tcgen05        (hm:952, hm:960)
//   3. recheck_kernel        pairs within `band` of gamma re-evaluated from the fp32 rows
//   4. greedy_scan_kernel    row i kept iff no kept j < i has bit (i, j)      (hm:958-961)
//   5. compact_kernel        kept bitmap -> ascending int64 row numbers       (hm:967)
//
// The N x N fp32 matrix of the reference (40 GB at N = 100k) is never formed; the bit
// matrix is N^2/8 bytes.
#include "common.cuh"
#include "sim_tc.cuh"
#include <vector>
#include <cstdio>
#include <cstdlib>

namespace hippo {

// ---- 3. exact re-evaluation of near-threshold pairs ---------------------------------
// One warp per pair.  Rows are normalised element-wise in fp32 exactly like hm:951
// (x / |x| with the fp32 norm), the products are accumulated in fp64, and the bit becomes
// !(sim < gamma) -- the best available stand-in for the reference's fp32 sgemm value.
__global__ void __launch_bounds__(256) recheck_kernel(const float* __restrict__ feats,
                                                      const float* __restrict__ norm, int d,
                                                      float gamma, const uint2* __restrict__ pairs,
                                                      const int32_t* __restrict__ count, int32_t cap,
                                                      uint32_t* __restrict__ mask, int64_t words_per_row) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int32_t np = *count;
  if (np > cap) np = cap;
  for (int64_t pi = warp; pi < np; pi += nwarps) {
    const uint2 pr = pairs[pi];
    const float* a = feats + (int64_t)pr.x * d;
    const float* b = feats + (int64_t)pr.y * d;
    const float na = norm[pr.x], nb = norm[pr.y];
    double acc = 0.0;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(a + c);
      const float4 y = *reinterpret_cast<const float4*>(b + c);
      acc += (double)__fdiv_rn(x.x, na) * (double)__fdiv_rn(y.x, nb);
      acc += (double)__fdiv_rn(x.y, na) * (double)__fdiv_rn(y.y, nb);
      acc += (double)__fdiv_rn(x.z, na) * (double)__fdiv_rn(y.z, nb);
      acc += (double)__fdiv_rn(x.w, na) * (double)__fdiv_rn(y.w, nb);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const bool bit = !(acc < (double)gamma);
      uint32_t* w = mask + (int64_t)pr.x * words_per_row + (pr.y >> 5);
      const uint32_t m = 1u << (pr.y & 31);
      if (bit) atomicOr(w, m); else atomicAnd(w, ~m);
    }
  }
}

// ---- 4. greedy scan over the bit matrix ------------------------------------------------
// CTA b owns rows [512 b, 512 b + 512), one thread per row.  The chain over blocks is sequential (a row's
// fate depends on which earlier rows were KEPT), so the kernel is bound by the hand-off latency between
// consecutive blocks; everything that does not depend on the predecessors' results is done ahead of them:
//   * kept-words are published as self-validating 64-bit values (tag << 32 | word), so a consumer polls
//     the data itself -- one L2 round trip per hand-off, no separate flag, no fence.  One warp per CTA
//     polls, and CTAs far behind the frontier sleep between polls: a line hammered by every waiting warp
//     of the grid delayed the publisher's store by ~5 us (profiles/);
//   * a thread holds its row's mask words against the block itself and its three predecessors in
//     registers; blocks further back are folded in from global memory as they are published (they are
//     final well before this block is on the critical path);
//   * the 512 x 512 diagonal block is resolved in ROUNDS over the whole block instead of a 512-step chain:
//     an undecided row is dropped as soon as a conflicting earlier row is known kept, and kept as soon as no
//     conflicting earlier row is still undecided (depth of the conflict chains, ~two rounds per kept row
//     of a scene on video-like data, one barrier per round);
//   * rows with no conflict bit against the three predecessors are resolved (phase A) BEFORE those
//     predecessors publish; only the rest -- typically the scene straddling the block boundary -- is
//     left for the critical path (phase B).
constexpr int kScanRows = 512;
constexpr int kScanWords = kScanRows / 32;   // 16
constexpr int kScanThreads = kScanRows;
constexpr int kScanPreds = 3;
constexpr unsigned long long kKeptTag = 1ull << 32;

__device__ __forceinline__ unsigned long long kept_poll(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
// the 16 mask words of `row` against the columns of 512-row block `blk` (64 contiguous bytes)
__device__ __forceinline__ void load_row_words(const uint32_t* __restrict__ mask, int64_t words_per_row,
                                               int64_t row, int blk, uint32_t (&out)[kScanWords]) {
  const uint4* p = reinterpret_cast<const uint4*>(mask + row * words_per_row + (int64_t)blk * kScanWords);
#pragma unroll
  for (int i = 0; i < kScanWords / 4; ++i) {
    const uint4 v = __ldg(p + i);
    out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
  }
}
// OR over w of (words[w] & set[w]) for 16 words of shared memory
__device__ __forceinline__ uint32_t and_any16(const uint32_t (&words)[kScanWords], const uint32_t* set, int nquads) {
  uint32_t h = 0;
#pragma unroll
  for (int i = 0; i < kScanWords / 4; ++i) {
    if (i < nquads) {
      const uint4 q = *reinterpret_cast<const uint4*>(set + 4 * i);
      h |= (words[4 * i] & q.x) | (words[4 * i + 1] & q.y) | (words[4 * i + 2] & q.z) | (words[4 * i + 3] & q.w);
    }
  }
  return h;
}

__global__ void __launch_bounds__(kScanThreads, 1) greedy_scan_kernel(const uint32_t* __restrict__ mask,
                                                                      int64_t words_per_row, int64_t n,
                                                                      unsigned long long* kept /*[blocks * 16], zeroed*/,
                                                                      unsigned long long* dbg) {
  __shared__ __align__(16) uint32_t s_kept[2][kScanWords];
  __shared__ __align__(16) uint32_t s_undec[2][kScanWords];
  __shared__ __align__(16) uint32_t s_kw[2][kScanWords];            // bulk ring
  __shared__ __align__(16) uint32_t s_pk[kScanPreds][kScanWords];   // predecessors' kept-words
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t row = (int64_t)b * kScanRows + tid;
  const bool live = row < n;
  auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  if (dbg && tid == 0) dbg[b * 8 + 0] = now();

  // this row against its own block (bits j < row only; later words were never written) and the three before
  uint32_t dg[kScanWords], pp[kScanPreds][kScanWords];
#pragma unroll
  for (int w = 0; w < kScanWords; ++w) {
    dg[w] = 0;
#pragma unroll
    for (int j = 0; j < kScanPreds; ++j) pp[j][w] = 0;
  }
  if (live) {
    load_row_words(mask, words_per_row, row, b, dg);
#pragma unroll
    for (int w = 0; w < kScanWords; ++w) {
      if (w > wid) dg[w] = 0;
      else if (w == wid) dg[w] &= (1u << lane) - 1u;
    }
#pragma unroll
    for (int j = 0; j < kScanPreds; ++j)
      if (b - 1 - j >= 0) load_row_words(mask, words_per_row, row, b - 1 - j, pp[j]);
  }
  uint32_t ext = 0;
#pragma unroll
  for (int w = 0; w < kScanWords; ++w)
#pragma unroll
    for (int j = 0; j < kScanPreds; ++j) ext |= pp[j][w];
  bool blocked = ext != 0;          // depends on a predecessor that has not published yet

  // blocks 0 .. b-4 from global memory, one block (16 kept-words) per step; warp 0 polls
  uint32_t hit = 0;
  const int nbulk = max(b - kScanPreds, 0);
  for (int g = 0; g < nbulk; ++g) {
    uint4 m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = make_uint4(0, 0, 0, 0);
    if (live) {
      const uint4* p = reinterpret_cast<const uint4*>(mask + row * words_per_row + (int64_t)g * kScanWords);
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = __ldg(p + i);
    }
    if (wid == 0 && lane < kScanWords) {
      unsigned long long v;
      while (!((v = kept_poll(&kept[(int64_t)g * kScanWords + lane])) >> 32))
        if (b - g > 6) __nanosleep(2000);
      s_kw[g & 1][lane] = (uint32_t)v;
    }
    __syncthreads();
    const uint32_t* kw = s_kw[g & 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 q = *reinterpret_cast<const uint4*>(kw + 4 * i);
      hit |= (m[i].x & q.x) | (m[i].y & q.y) | (m[i].z & q.z) | (m[i].w & q.w);
    }
  }
  if (dbg && tid == 0) dbg[b * 8 + 1] = now();

  bool undec = live && hit == 0;
  bool kept_me = false;
  int round = 0;
  const int nquads = wid / 4 + 1;     // dg words beyond the row's own are zero
  // rounds until nothing changes; blocked rows stay undecided (and count as such for the others)
  auto run_rounds = [&]() {
    bool changed = true;
    for (;; ++round) {
      const int buf = round & 1;
      const uint32_t uw = __ballot_sync(0xffffffffu, undec);
      const uint32_t kwd = __ballot_sync(0xffffffffu, kept_me);
      if (lane == 0) { s_undec[buf][wid] = uw; s_kept[buf][wid] = kwd; }
      if (!__syncthreads_or(changed ? 1 : 0)) break;     // barrier + "did the last round decide anything?"
      changed = false;
      if (undec && !blocked) {
        if (and_any16(dg, s_kept[buf], nquads) != 0) { undec = false; changed = true; }              // a conflicting earlier row is kept
        else if (and_any16(dg, s_undec[buf], nquads) == 0) { undec = false; kept_me = true; changed = true; }  // none can still be kept
      }
    }
    ++round;   // the buffer written last stays intact for the next call's first barrier
  };
  run_rounds();                                       // phase A
  if (dbg && tid == 0) { dbg[b * 8 + 5] = now(); dbg[b * 8 + 6] = round; }

  // phase B: the three predecessors' kept-words (b-3 and b-2 by warp 0, b-1 by warp 1)
  if (b >= 1) {
    if (wid < 2) {
      const int j = wid == 0 ? 2 - (lane >> 4) : 0;   // predecessor b-1-j
      if ((wid == 0 || lane < kScanWords) && b - 1 - j >= 0) {
        unsigned long long v;
        while (!((v = kept_poll(&kept[(int64_t)(b - 1 - j) * kScanWords + (lane & 15)])) >> 32)) {}
        s_pk[j][lane & 15] = (uint32_t)v;
      } else if (wid == 0 || lane < kScanWords) {
        s_pk[j][lane & 15] = 0;
      }
    }
    __syncthreads();
    if (dbg && tid == 0) dbg[b * 8 + 2] = now();
    if (blocked) {
      uint32_t h = 0;
#pragma unroll
      for (int j = 0; j < kScanPreds; ++j) h |= and_any16(pp[j], s_pk[j], 4);
      if (h != 0) undec = false;
      blocked = false;
    }
    run_rounds();
  } else if (dbg && tid == 0) dbg[b * 8 + 2] = now();

  {
    const uint32_t kwd = __ballot_sync(0xffffffffu, kept_me);
    if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&kept[(int64_t)b * kScanWords + wid]) = kKeptTag | kwd;
    if (dbg && tid == 0) { dbg[b * 8 + 3] = now(); dbg[b * 8 + 4] = round; }
  }
}

// ---- 5. kept bitmap -> ascending row numbers ---------------------------------------------
__global__ void __launch_bounds__(1024) compact_kernel(const unsigned long long* __restrict__ kept, int64_t nwords,
                                                       int64_t n, int64_t* __restrict__ out_keep,
                                                       int32_t* __restrict__ out_count) {
  __shared__ int64_t s_sum[1024];
  const int t = threadIdx.x;
  const int64_t per = (nwords + 1023) / 1024;
  const int64_t w0 = t * per, w1 = min(w0 + per, nwords);
  int64_t c = 0;
  for (int64_t w = w0; w < w1; ++w) c += __popc((uint32_t)kept[w]);
  s_sum[t] = c;
  __syncthreads();
  // inclusive scan (Hillis-Steele; 1024 entries, one launch per consolidation)
  for (int o = 1; o < 1024; o <<= 1) {
    int64_t v = t >= o ? s_sum[t - o] : 0;
    __syncthreads();
    s_sum[t] += v;
    __syncthreads();
  }
  int64_t pos = s_sum[t] - c;
  for (int64_t w = w0; w < w1; ++w) {
    uint32_t bits = (uint32_t)kept[w];   // low half of the tagged word
    while (bits) {
      const int bpos = __ffs(bits) - 1;
      bits &= bits - 1;
      const int64_t row = w * 32 + bpos;
      if (row < n) out_keep[pos++] = row;
    }
  }
  if (t == 1023) *out_count = (int32_t)s_sum[1023];
}

__global__ void iota_kernel(int64_t n, int64_t* out_keep, int32_t* out_count) {
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out_keep[i] = i;
  if (threadIdx.x == 0) *out_count = (int32_t)n;
}

__global__ void stats_kernel(const int32_t* unc_count, int32_t cap, const int32_t* inexact, int32_t* out_stats) {
  const int32_t c = *unc_count;
  out_stats[0] = c < cap ? c : cap;
  out_stats[1] = c > cap ? 1 : 0;
  out_stats[2] = *inexact;
  out_stats[3] = 0;
}

struct ConsLayout {
  __nv_bfloat16* bf;
  float* norm;
  uint32_t* mask;
  int64_t words_per_row;
  unsigned long long* kept;   // tagged kept-words (greedy_scan_kernel)
  int64_t kept_words;
  uint2* unc;
  int32_t unc_cap;
  int32_t* counters;   // [0] uncertain count, [1] ready, [2] inexact
  size_t bytes;
};

static ConsLayout cons_layout(void* ws, size_t ws_bytes, int64_t n, int d) {
  Carver c(ws, ws_bytes);
  ConsLayout L{};
  L.bf = c.take<__nv_bfloat16>((size_t)n * d);
  L.norm = c.take<float>((size_t)n);
  L.words_per_row = (n + kTcBN - 1) / kTcBN * (kTcBN / 32);
  L.mask = c.take<uint32_t>((size_t)n * L.words_per_row);
  L.kept_words = (n + kScanRows - 1) / kScanRows * kScanWords;
  L.kept = c.take<unsigned long long>((size_t)L.kept_words);
  int64_t cap = 64 * n + (1 << 20);
  if (cap > 0x3fffffff) cap = 0x3fffffff;
  L.unc_cap = (int32_t)cap;
  L.unc = c.take<uint2>((size_t)cap);
  L.counters = c.take<int32_t>(64);
  L.bytes = c.used();
  return L;
}

}  // namespace hippo

extern "C" {

size_t hippo_consolidate_workspace_bytes(int64_t n, int32_t d) {
  if (n <= 2 || d <= 0) return 256;
  return hippo::cons_layout(nullptr, 0, n, d).bytes;
}

hippo_status hippo_consolidate(const float* feats, int64_t n, int32_t d, float gamma, float band_exact,
                               float band_inexact, int64_t* out_keep, int32_t* out_count,
                               int32_t* out_stats, void* ws, size_t ws_bytes, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(n >= 0 && d > 0 && d % 64 == 0, "hippo_consolidate: need d %% 64 == 0 (d=%d)", d);
  HIPPO_REQUIRE(n < 0x7fffffffll, "hippo_consolidate: n too large");
  HIPPO_REQUIRE(out_count != nullptr && (n == 0 || (feats && out_keep)), "hippo_consolidate: null pointer");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 2) {  // hm:946-947
    iota_kernel<<<1, 32, 0, s>>>(n, out_keep, out_count);
    if (out_stats) HIPPO_CUDA(cudaMemsetAsync(out_stats, 0, 16, s));
    HIPPO_CUDA(cudaGetLastError());
    return HIPPO_OK;
  }
  ConsLayout L = cons_layout(ws, ws_bytes, n, d);
  if (ws == nullptr || ((uintptr_t)ws & 255) || L.bytes > ws_bytes) {
    set_error("hippo_consolidate: workspace of %zu bytes needed (256-byte aligned), got %zu", L.bytes, ws_bytes);
    return HIPPO_E_WORKSPACE;
  }
  HIPPO_CUDA(cudaMemsetAsync(L.counters, 0, 64 * sizeof(int32_t), s));
  st = hippo_bank_build(feats, HIPPO_F32, n, d, d, L.bf, L.norm, L.counters + 2, stream);
  if (st != HIPPO_OK) return st;

  TcMaskArgs a{};
  a.feats_bf16 = L.bf;
  a.norm = L.norm;
  a.n = n;
  a.d = d;
  a.gamma = gamma;
  a.band_exact = band_exact;
  a.band_inexact = band_inexact;
  a.inexact = L.counters + 2;
  a.mask = L.mask;
  a.words_per_row = L.words_per_row;
  a.uncertain = L.unc;
  a.uncertain_count = L.counters + 0;
  a.uncertain_cap = L.unc_cap;
  st = tc_mask_launch(a, s);
  if (st != HIPPO_OK) return st;

  recheck_kernel<<<sm_count() * 8, 256, 0, s>>>(feats, L.norm, d, gamma, L.unc, L.counters + 0, L.unc_cap,
                                                 L.mask, L.words_per_row);
  HIPPO_CUDA(cudaGetLastError());
  const int scan_blocks = (int)((n + kScanRows - 1) / kScanRows);
  HIPPO_CUDA(cudaMemsetAsync(L.kept, 0, (size_t)L.kept_words * sizeof(unsigned long long), s));
  unsigned long long* dbg = nullptr;
  if (getenv("HIPPO_SCAN_DEBUG")) cudaMalloc(&dbg, (size_t)scan_blocks * 64);
  greedy_scan_kernel<<<scan_blocks, kScanThreads, 0, s>>>(L.mask, L.words_per_row, n, L.kept, dbg);
  if (dbg) {
    cudaStreamSynchronize(s);
    std::vector<unsigned long long> h((size_t)scan_blocks * 8);
    cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(dbg);
    const unsigned long long t0 = h[0];
    for (int i = 0; i < scan_blocks; ++i)
      if (i < 6 || i % 16 == 0 || i > scan_blocks - 4)
        fprintf(stderr, "[scan] blk %3d start %8.2f bulk_done %8.2f phaseA %8.2f (%llu rounds) preds %8.2f published %8.2f us  (rounds %llu, since prev publish %.2f)\n", i,
                (h[i * 8] - t0) / 1e3, (h[i * 8 + 1] - t0) / 1e3, (h[i * 8 + 5] - t0) / 1e3, h[i * 8 + 6], (h[i * 8 + 2] - t0) / 1e3, (h[i * 8 + 3] - t0) / 1e3, h[i * 8 + 4],
                i ? (h[i * 8 + 3] - h[(i - 1) * 8 + 3]) / 1e3 : 0.0);
  }
  HIPPO_CUDA(cudaGetLastError());
  compact_kernel<<<1, 1024, 0, s>>>(L.kept, L.kept_words, n, out_keep, out_count);
  HIPPO_CUDA(cudaGetLastError());
  if (out_stats) {
    stats_kernel<<<1, 1, 0, s>>>(L.counters + 0, L.unc_cap, L.counters + 2, out_stats);
    HIPPO_CUDA(cudaGetLastError());
  }
  return HIPPO_OK;
}

}  // extern "C"
