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
// CTA b owns rows [512 b, 512 b + 512).  The chain over blocks is sequential (a row's fate depends on
// which earlier rows were KEPT), so the kernel is bound by the hand-off latency between consecutive
// blocks; everything that does not depend on the predecessor's result is done ahead of it:
//   * kept-words are published as self-validating 64-bit values (tag << 32 | word), so a consumer polls
//     the data itself -- one L2 round trip per hand-off, no separate flag, no fence;
//   * the mask columns of the two preceding blocks and the 512 x 512 diagonal block are staged in
//     shared memory before their kept-words exist; blocks further back are folded in from global
//     memory as they are published (they are final long before this block is on the critical path);
//   * the diagonal block is resolved by one warp, 32 rows at a time, not by a 32-step chain but by
//     rounds of two ballots: every candidate with no surviving earlier conflicting candidate is
//     kept at once, everything those rows suppress is dropped (1-2 rounds on video-like data).
constexpr int kScanRows = 512;
constexpr int kScanWords = kScanRows / 32;   // 16
constexpr int kScanThreads = 512;
constexpr int kScanStride = kScanWords + 1;  // 17: conflict-free rows
constexpr int kScanPrev = 2;                 // predecessor blocks staged in shared memory
constexpr size_t kScanSmem = (size_t)(1 + kScanPrev) * kScanRows * kScanStride * sizeof(uint32_t);
constexpr unsigned long long kKeptTag = 1ull << 32;

__device__ __forceinline__ unsigned long long kept_poll(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

__global__ void __launch_bounds__(kScanThreads, 2) greedy_scan_kernel(const uint32_t* __restrict__ mask,
                                                                   int64_t words_per_row, int64_t n,
                                                                   unsigned long long* kept /*[blocks * 16], zeroed*/) {
  extern __shared__ uint32_t s_scan[];
  uint32_t* s_diag = s_scan;                                     // [512][17]
  uint32_t* s_prev = s_scan + kScanRows * kScanStride;           // [2][512][17]: columns of block b-1, b-2
  __shared__ uint32_t s_kw[kScanPrev][kScanWords];
  __shared__ uint32_t s_supp[kScanRows];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;   // 16 warps x 32 rows
  const int64_t r0 = (int64_t)b * kScanRows;

  // diagonal block and the columns of the two preceding blocks -> smem (rows beyond n: zero bits)
  for (int i = tid; i < kScanRows * kScanWords; i += kScanThreads) {
    const int r = i / kScanWords, w = i - r * kScanWords;
    const int64_t row = r0 + r;
    uint32_t v = 0;
    // words beyond the row's own position were never written by the contraction
    if (row < n && (int64_t)b * kScanWords + w <= (row >> 5)) v = __ldg(&mask[row * words_per_row + (int64_t)b * kScanWords + w]);
    s_diag[r * kScanStride + w] = v;
#pragma unroll
    for (int pb = 1; pb <= kScanPrev; ++pb) {
      uint32_t u = 0;
      if (row < n && b - pb >= 0) u = __ldg(&mask[row * words_per_row + (int64_t)(b - pb) * kScanWords + w]);
      s_prev[((pb - 1) * kScanRows + r) * kScanStride + w] = u;
    }
  }

  // blocks 0 .. b-3 from global memory: warp = 32 rows, lane = one kept-word of a group of 32
  uint32_t pre[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) pre[r] = 0;
  const int64_t nbulk = (int64_t)max(b - kScanPrev, 0) * kScanWords;
  for (int64_t w0 = 0; w0 < nbulk; w0 += 32) {
    const int64_t w = w0 + lane;
    const bool in = w < nbulk;
    unsigned long long v = 0;
    do {
      if (in && !(v >> 32)) v = kept_poll(&kept[w]);
    } while (!__all_sync(0xffffffffu, !in || (v >> 32) != 0));
    const uint32_t kw = (uint32_t)v;
    if (kw != 0) {
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const int64_t row = r0 + wid * 32 + r;
        if (row < n) pre[r] |= __ldg(&mask[row * words_per_row + w]) & kw;
      }
    }
  }
  bool supp = r0 + tid >= n;       // thread = row from here on
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const unsigned any = __ballot_sync(0xffffffffu, pre[r] != 0);
    if (lane == r) supp |= any != 0;
  }

  // the two predecessors: poll their 16 + 16 kept-words (b-2 is published before b-1, waiting for both loses nothing)
  if (wid == 0) {
    const int pb = 1 + (lane >> 4), w = lane & 15;
    uint32_t kw = 0;
    if (b - pb >= 0) {
      unsigned long long v;
      do { v = kept_poll(&kept[(int64_t)(b - pb) * kScanWords + w]); } while (!(v >> 32));
      kw = (uint32_t)v;
    }
    s_kw[pb - 1][w] = kw;
  }
  __syncthreads();                 // also orders the smem staging above
  {
    uint32_t hit = 0;
#pragma unroll
    for (int pb = 0; pb < kScanPrev; ++pb)
#pragma unroll
      for (int w = 0; w < kScanWords; ++w) hit |= s_prev[(pb * kScanRows + tid) * kScanStride + w] & s_kw[pb][w];
    s_supp[tid] = (supp || hit != 0) ? 1u : 0u;
  }
  __syncthreads();

  if (wid == 0) {
    // lane l carries rows sb * 32 + l of every sub-block; bit sb of supp16 = that row is suppressed
    uint32_t supp16 = 0;
#pragma unroll
    for (int sb = 0; sb < kScanWords; ++sb) supp16 |= s_supp[sb * 32 + lane] << sb;
#pragma unroll
    for (int sb = 0; sb < kScanWords; ++sb) {
      // words of the LATER sub-blocks against this one (independent of the outcome: loaded first)
      uint32_t upd[kScanWords];
#pragma unroll
      for (int s2 = sb + 1; s2 < kScanWords; ++s2) upd[s2] = s_diag[(s2 * 32 + lane) * kScanStride + sb];
      const uint32_t dw = s_diag[(sb * 32 + lane) * kScanStride + sb];    // bits j < lane only
      uint32_t cand = __ballot_sync(0xffffffffu, !((supp16 >> sb) & 1u));
      uint32_t km = 0;
      while (cand) {
        const bool mine = (cand >> lane) & 1u;
        const uint32_t keep_now = __ballot_sync(0xffffffffu, mine && (dw & cand) == 0);   // never empty: the lowest candidate
        km |= keep_now;
        const uint32_t killed = __ballot_sync(0xffffffffu, mine && (dw & keep_now) != 0);
        cand &= ~(keep_now | killed);
      }
      if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&kept[(int64_t)b * kScanWords + sb]) = kKeptTag | km;
#pragma unroll
      for (int s2 = sb + 1; s2 < kScanWords; ++s2) supp16 |= ((upd[s2] & km) != 0 ? 1u : 0u) << s2;
    }
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
  HIPPO_CUDA(cudaFuncSetAttribute(greedy_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmem));
  greedy_scan_kernel<<<scan_blocks, kScanThreads, kScanSmem, s>>>(L.mask, L.words_per_row, n, L.kept);
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
