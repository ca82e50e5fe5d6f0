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
// CTA b owns rows [512 b, 512 b + 512).  It ANDs its rows' mask words against the final
// kept-words of earlier blocks as those become final (ordered chain through `ready`), then
// one warp resolves the 512 x 512 diagonal block 32 rows at a time.
constexpr int kScanRows = 512;
constexpr int kScanWords = kScanRows / 32;   // 16
constexpr int kScanThreads = 512;

__global__ void __launch_bounds__(kScanThreads) greedy_scan_kernel(const uint32_t* __restrict__ mask,
                                                                   int64_t words_per_row, int64_t n,
                                                                   uint32_t* kept /*[ceil(n/32)] padded to blocks*/,
                                                                   int32_t* ready) {
  __shared__ uint32_t s_diag[kScanRows][kScanWords + 1];
  __shared__ uint32_t s_pre[kScanRows];
  __shared__ int s_avail;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;   // 16 warps x 32 rows
  const int64_t r0 = (int64_t)b * kScanRows;

  // diagonal block -> smem (rows beyond n: zero bits, treated as suppressed later)
  for (int i = threadIdx.x; i < kScanRows * kScanWords; i += kScanThreads) {
    const int r = i / kScanWords, w = i - r * kScanWords;
    const int64_t row = r0 + r;
    uint32_t v = 0;
    // words beyond the row's own position were never written by the contraction
    if (row < n && (int64_t)b * kScanWords + w <= (row >> 5)) v = mask[row * words_per_row + (int64_t)b * kScanWords + w];
    s_diag[r][w] = v;
  }
  uint32_t pre[32 / 1];  // per warp: 32 rows; lane l accumulates for every row, reduced by ballot
#pragma unroll
  for (int r = 0; r < 32; ++r) pre[r] = 0;

  int done = 0;
  while (done < b) {
    if (threadIdx.x == 0) {
      int a;
      do {
        a = *((volatile int32_t*)ready);
      } while (a <= done);
      s_avail = a < b ? a : b;
    }
    __syncthreads();
    const int avail = s_avail;
    __threadfence();   // acquire side of the kept-words published before `ready` moved
    const int64_t wlo = (int64_t)done * kScanWords, whi = (int64_t)avail * kScanWords;
    for (int64_t w = wlo + lane; w < whi; w += 32) {
      const uint32_t kw = __ldcg(&kept[w]);
      if (kw != 0) {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const int64_t row = r0 + wid * 32 + r;
          if (row < n) pre[r] |= __ldg(&mask[row * words_per_row + w]) & kw;
        }
      }
    }
    done = avail;
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const unsigned any = __ballot_sync(0xffffffffu, pre[r] != 0);
    if (lane == 0) s_pre[wid * 32 + r] = any;
  }
  __syncthreads();

  if (wid == 0) {
    uint32_t keptloc[kScanWords];
#pragma unroll
    for (int sb = 0; sb < kScanWords; ++sb) {
      const int r = sb * 32 + lane;
      const int64_t row = r0 + r;
      bool supp = (row >= n) || (s_pre[r] != 0);
#pragma unroll
      for (int w = 0; w < kScanWords; ++w)
        if (w < sb) supp |= (s_diag[r][w] & keptloc[w]) != 0;
      const uint32_t dw = s_diag[r][sb];
      const uint32_t suppmask = __ballot_sync(0xffffffffu, supp);
      uint32_t km = 0;
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        const uint32_t wl = __shfl_sync(0xffffffffu, dw, l);
        const bool keep = ((wl & km) == 0) && !((suppmask >> l) & 1u);
        km |= keep ? (1u << l) : 0u;
      }
      keptloc[sb] = km;
      if (lane == 0) kept[(int64_t)b * kScanWords + sb] = km;
    }
    __threadfence();
    if (lane == 0) atomicExch(ready, b + 1);
  }
}

// ---- 5. kept bitmap -> ascending row numbers ---------------------------------------------
__global__ void __launch_bounds__(1024) compact_kernel(const uint32_t* __restrict__ kept, int64_t nwords,
                                                       int64_t n, int64_t* __restrict__ out_keep,
                                                       int32_t* __restrict__ out_count) {
  __shared__ int64_t s_sum[1024];
  const int t = threadIdx.x;
  const int64_t per = (nwords + 1023) / 1024;
  const int64_t w0 = t * per, w1 = min(w0 + per, nwords);
  int64_t c = 0;
  for (int64_t w = w0; w < w1; ++w) c += __popc(kept[w]);
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
    uint32_t bits = kept[w];
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
  uint32_t* kept;
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
  L.kept = c.take<uint32_t>((size_t)L.kept_words);
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
  greedy_scan_kernel<<<scan_blocks, kScanThreads, 0, s>>>(L.mask, L.words_per_row, n, L.kept, L.counters + 1);
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
