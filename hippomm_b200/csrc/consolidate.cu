// hippo_consolidate: greedy cosine-redundancy filter (reference: _select_key_frames, hm:944-967).
//
// The reference forms the whole N x N similarity matrix (hm:952) and then walks the rows: row i is kept iff
// every KEPT earlier row j has sim(i, j) < gamma (hm:958-961).  Similarities against rows that were dropped
// are never looked at -- and on video-like input most rows are dropped.  So the rows are processed in BANDS
// of kConsBand rows, and a band is only contracted against
//     (a) the rows kept so far, held compacted at the front of a second bf16 matrix Y, and
//     (b) itself (lower triangle),
// which is exactly the set of pairs the greedy rule can consult.  With K rows kept out of N the tensor work
// drops from N^2/2 to about N K / 2 + N band / 2 pairs (10x at 6% kept); when everything is kept it is the
// full triangle again.  The decisions are identical either way.
//
//   0. hippo_bank_build       fp32 rows -> bf16 rows X + fp32 norms                (hm:951)
//   per band (all asynchronous, the extent of a band's work is read from device memory):
//   1. cons_advance_kernel    finishes the previous band (its kept rows are compacted to Y[K ..), K grows,
//                             the caller's out_keep receives their row numbers) and copies this band's rows
//                             to Y[K .. K + rows)
//   2. sim_tc EPI_MASK        similarity bits of Y rows [K, K + rows) against Y rows before them, on tcgen05
//                                                                                 (hm:952, hm:960)
//   3. recheck_kernel         pairs within `band` of gamma re-evaluated from the fp32 rows
//   4. greedy_scan_kernel     row i kept iff no kept j < i has bit (i, j)           (hm:958-961)
//
// The N x N fp32 matrix of the reference (40 GB at N = 100k) is never formed; the bit matrix holds the rows of
// the current band only ((band + 1024) x N / 8 bytes: 115 MB at N = 100k).
#include "common.cuh"
#include "sim_tc.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace hippo {

constexpr int kConsBandDefault = 8192;

// ---- 3. exact re-evaluation of near-threshold pairs ---------------------------------
// One warp per pair (i, j) of Y rows.  The rows are looked up in the caller's fp32 matrix through yidx,
// normalised element-wise in fp32 exactly like hm:951 (x / |x| with the fp32 norm), the products are
// accumulated in fp64, and the bit becomes !(sim < gamma) -- the best available stand-in for the
// reference's fp32 sgemm value.
__global__ void __launch_bounds__(256) recheck_kernel(const float* __restrict__ feats,
                                                      const float* __restrict__ norm, int d,
                                                      const int64_t* __restrict__ yidx,
                                                      float gamma, const uint2* __restrict__ pairs,
                                                      const int32_t* __restrict__ count, int32_t cap,
                                                      uint32_t* __restrict__ mask, int64_t words_per_row,
                                                      const int32_t* __restrict__ dyn_k) {
  const int64_t row0 = (int64_t)(*dyn_k / 512) * 512;     // first row held by the band-local bit matrix
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int32_t np = *count;
  if (np > cap) np = cap;
  for (int64_t pi = warp; pi < np; pi += nwarps) {
    const uint2 pr = pairs[pi];
    const int64_t ra = yidx[pr.x], rb = yidx[pr.y];
    const float* a = feats + ra * d;
    const float* b = feats + rb * d;
    const float na = norm[ra], nb = norm[rb];
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
      uint32_t* w = mask + ((int64_t)pr.x - row0) * words_per_row + (pr.y >> 5);
      const uint32_t m = 1u << (pr.y & 31);
      if (bit) atomicOr(w, m); else atomicAnd(w, ~m);
    }
  }
}

// ---- 4. greedy scan over the bit matrix of one band --------------------------------------
// Rows are numbered in Y space: rows [0, K) are final (kept), the band is [K, n), n = K + band_rows.  CTA x
// owns the 512 rows of block b = K / 512 + x, one thread per row (rows of the first block that lie below K
// are simply resolved again: kept rows never conflict with each other).  The chain over blocks is sequential
// (a row's fate depends on which earlier rows were KEPT), so the kernel is bound by the hand-off latency
// between consecutive blocks; everything that does not depend on the predecessors' results is done ahead:
//   * kept-words are published as self-validating 64-bit values (tag << 32 | word), so a consumer polls
//     the data itself -- one L2 round trip per hand-off, no separate flag, no fence.  One warp per CTA
//     polls, and CTAs far behind the frontier sleep between polls: a line hammered by every waiting warp
//     of the grid delayed the publisher's store by ~5 us (profiles/).  Blocks below K need no polling:
//     all their rows are kept;
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
                                                                      int64_t words_per_row,
                                                                      const int32_t* __restrict__ dyn_k, int band_rows,
                                                                      unsigned long long* kept /*[grid * 16], zeroed*/,
                                                                      unsigned long long* dbg) {
  __shared__ __align__(16) uint32_t s_kept[2][kScanWords];
  __shared__ __align__(16) uint32_t s_undec[2][kScanWords];
  __shared__ __align__(16) uint32_t s_kw[2][kScanWords];            // bulk ring
  __shared__ __align__(16) uint32_t s_pk[kScanPreds][kScanWords];   // predecessors' kept-words
  const int kfinal = *dyn_k;
  const int64_t n = (int64_t)kfinal + band_rows;
  const int b0 = kfinal / kScanRows;               // first block of this launch; blocks below it are all kept
  const int b = b0 + blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t row = (int64_t)blockIdx.x * kScanRows + tid;   // row of the band-local bit matrix (Y row - 512 b0)
  if ((int64_t)b * kScanRows >= n) return;         // nobody waits for a block past the end
  const bool live = (int64_t)b * kScanRows + tid < n;
  auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  if (dbg && tid == 0) dbg[blockIdx.x * 8 + 0] = now();

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
  // a predecessor below b0 is already known (all kept): only predecessors of this launch can block a row
  bool blocked = false;
#pragma unroll
  for (int j = 0; j < kScanPreds; ++j) {
    uint32_t e = 0;
#pragma unroll
    for (int w = 0; w < kScanWords; ++w) e |= pp[j][w];
    if (b - 1 - j >= b0 && e != 0) blocked = true;
  }

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
    if (g < b0) {                                  // rows of a finished band: all kept, nothing to wait for
      hit |= (m[0].x | m[0].y | m[0].z | m[0].w) | (m[1].x | m[1].y | m[1].z | m[1].w) |
             (m[2].x | m[2].y | m[2].z | m[2].w) | (m[3].x | m[3].y | m[3].z | m[3].w);
      continue;
    }
    if (wid == 0 && lane < kScanWords) {
      unsigned long long v;
      while (!((v = kept_poll(&kept[(int64_t)(g - b0) * kScanWords + lane])) >> 32))
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
  // predecessors that belong to finished bands need no hand-off either
#pragma unroll
  for (int j = 0; j < kScanPreds; ++j) {
    if (b - 1 - j >= 0 && b - 1 - j < b0) {
#pragma unroll
      for (int w = 0; w < kScanWords; ++w) hit |= pp[j][w];
    }
  }
  if (dbg && tid == 0) dbg[blockIdx.x * 8 + 1] = now();

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
  if (dbg && tid == 0) { dbg[blockIdx.x * 8 + 5] = now(); dbg[blockIdx.x * 8 + 6] = round; }

  // phase B: the kept-words of the predecessors that belong to this launch (b-3 and b-2 by warp 0, b-1 by warp 1)
  if (b - 1 >= b0) {
    if (wid < 2) {
      const int j = wid == 0 ? 2 - (lane >> 4) : 0;   // predecessor b-1-j
      if (wid == 0 || lane < kScanWords) {
        uint32_t kwv = 0;
        if (b - 1 - j >= b0) {
          unsigned long long v;
          while (!((v = kept_poll(&kept[(int64_t)(b - 1 - j - b0) * kScanWords + (lane & 15)])) >> 32)) {}
          kwv = (uint32_t)v;
        }
        s_pk[j][lane & 15] = kwv;                     // finished-band predecessors were folded in above
      }
    }
    __syncthreads();
    if (dbg && tid == 0) dbg[blockIdx.x * 8 + 2] = now();
    if (blocked) {
      uint32_t h = 0;
#pragma unroll
      for (int j = 0; j < kScanPreds; ++j) h |= and_any16(pp[j], s_pk[j], 4);
      if (h != 0) undec = false;
      blocked = false;
    }
    run_rounds();
  } else if (dbg && tid == 0) dbg[blockIdx.x * 8 + 2] = now();

  {
    const uint32_t kwd = __ballot_sync(0xffffffffu, kept_me);
    if (lane == 0)
      *reinterpret_cast<volatile unsigned long long*>(&kept[(int64_t)blockIdx.x * kScanWords + wid]) = kKeptTag | kwd;
    if (dbg && tid == 0) { dbg[blockIdx.x * 8 + 3] = now(); dbg[blockIdx.x * 8 + 4] = round; }
  }
}

// ---- 1. finish the previous band, stage the next one -----------------------------------------
// dyn[par_in] = K before the previous band, whose rows sit at Y[K, K + prev_rows) (original rows prev_r0 ..)
// and whose kept-words are kept_prev (block-local: bit position = Y row - 512 * (K / 512)).  The kept rows
// are compacted to Y[K, K'), out_keep[K ..) receives their original row numbers (ascending: hm:967), K' goes
// to dyn[par_in ^ 1], and the next band's rows [next_r0, next_r0 + next_rows) are copied to Y[K', ...).
// Everything is read from X (the bf16 image of the caller's rows), never from Y, so the compaction cannot
// trample rows it still needs.  Every CTA recomputes the (short) prefix over the kept-words on its own.
// one bf16 row, 16 bytes per lane and step, four loads in flight per lane
__device__ __forceinline__ void copy_row(const uint4* __restrict__ sp, uint4* __restrict__ dp, int nvec, int lane) {
  int v = lane;
  for (; v + 96 < nvec; v += 128) {
    const uint4 a = ldg_stream(sp + v), b = ldg_stream(sp + v + 32), c = ldg_stream(sp + v + 64), e = ldg_stream(sp + v + 96);
    dp[v] = a; dp[v + 32] = b; dp[v + 64] = c; dp[v + 96] = e;
  }
  for (; v < nvec; v += 32) dp[v] = ldg_stream(sp + v);
}

constexpr int kAdvThreads = 256;
constexpr int kAdvMaxWords = 1024;   // band of at most 32k - 512 rows

__global__ void __launch_bounds__(kAdvThreads) cons_advance_kernel(
    const __nv_bfloat16* __restrict__ X, const float* __restrict__ xnorm, int d, __nv_bfloat16* __restrict__ Y,
    float* __restrict__ ynorm, int64_t* __restrict__ yidx, const unsigned long long* __restrict__ kept_prev,
    unsigned long long* __restrict__ kept_next, int kept_words, int32_t* dyn, int par_in, int64_t prev_r0,
    int prev_rows, int64_t next_r0, int next_rows, int32_t* unc_count, int32_t unc_cap, int32_t* stats) {
  __shared__ int s_pref[kAdvMaxWords + 1];
  __shared__ uint32_t s_bits[kAdvMaxWords];
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = dyn[par_in];
  const int off = K - (K / kScanRows) * kScanRows;        // block-local position of Y row K
  const int nwords = prev_rows > 0 ? (off + prev_rows + 31) / 32 : 0;
  for (int w = tid; w < nwords; w += kAdvThreads) {
    uint32_t bits = (uint32_t)kept_prev[w];
    if (w == off / 32) bits &= ~((1u << (off & 31)) - 1u);            // rows below K are old
    if (w < off / 32) bits = 0;
    const int end = off + prev_rows - w * 32;                          // rows past the band
    if (end < 32) bits &= (1u << end) - 1u;
    s_bits[w] = bits;
  }
  __syncthreads();
  if (tid < 32) {      // exclusive prefix of the popcounts: one warp, 32 words per step
    int carry = 0;
    for (int w0 = 0; w0 < nwords; w0 += 32) {
      const int w = w0 + tid;
      const int c = w < nwords ? __popc(s_bits[w]) : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += v; }
      if (w < nwords) s_pref[w] = carry + incl - c;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (tid == 0) s_pref[nwords] = carry;
  }
  __syncthreads();
  const int added = s_pref[nwords];
  const int K2 = K + added;

  const int64_t warp = (int64_t)blockIdx.x * (kAdvThreads / 32) + (tid >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kAdvThreads / 32);
  const int vec_per_row = d / 8;                                       // 16-byte vectors per bf16 row
  // compaction: one warp per kept row
  for (int64_t t = warp; t < prev_rows; t += nwarps) {
    const int pos = off + (int)t;
    const uint32_t bits = s_bits[pos >> 5];
    if (!((bits >> (pos & 31)) & 1u)) continue;
    const int rank = s_pref[pos >> 5] + __popc(bits & ((1u << (pos & 31)) - 1u));
    const int64_t src = prev_r0 + t, dst = (int64_t)K + rank;
    copy_row(reinterpret_cast<const uint4*>(X + src * d), reinterpret_cast<uint4*>(Y + dst * d), vec_per_row, lane);
    if (lane == 0) { ynorm[dst] = xnorm[src]; yidx[dst] = src; }
  }
  // next band
  for (int64_t t = warp; t < next_rows; t += nwarps) {
    const int64_t src = next_r0 + t, dst = (int64_t)K2 + t;
    copy_row(reinterpret_cast<const uint4*>(X + src * d), reinterpret_cast<uint4*>(Y + dst * d), vec_per_row, lane);
    if (lane == 0) { ynorm[dst] = xnorm[src]; yidx[dst] = src; }
  }
  // the next band's hand-off words, the uncertain-pair list, the running statistics
  for (int64_t i = (int64_t)blockIdx.x * kAdvThreads + tid; i < kept_words; i += (int64_t)gridDim.x * kAdvThreads)
    kept_next[i] = 0;
  if (blockIdx.x == 0 && tid == 0) {
    dyn[par_in ^ 1] = K2;
    const int32_t c = *unc_count;
    stats[0] += c < unc_cap ? c : unc_cap;
    if (c > unc_cap) stats[1] = 1;
    *unc_count = 0;
  }
}

__global__ void cons_finish_kernel(const int32_t* dyn, int par, const int32_t* inexact, const int32_t* stats,
                                   int32_t* out_count, int32_t* out_stats) {
  *out_count = dyn[par];
  if (out_stats) { out_stats[0] = stats[0]; out_stats[1] = stats[1]; out_stats[2] = *inexact; out_stats[3] = 0; }
}

__global__ void iota_kernel(int64_t n, int64_t* out_keep, int32_t* out_count) {
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out_keep[i] = i;
  if (threadIdx.x == 0) *out_count = (int32_t)n;
}

// HIPPO_CONS_TIMING: CUDA-event time of every launch of the last hippo_consolidate call on this thread, summed per
// stage {mask (tcgen05), recheck, scan, advance / staging / bank build}; read back with hippo_debug_consolidate_timing
static thread_local double g_cons_ms[4] = {0, 0, 0, 0};

static int cons_band(int requested = 0) {
  const char* e = getenv("HIPPO_CONS_BAND");
  int b = requested > 0 ? requested : (e ? atoi(e) : kConsBandDefault);
  if (b < kScanRows) b = kScanRows;
  if (b > (kAdvMaxWords - 32) * 32) b = (kAdvMaxWords - 32) * 32;
  return b / kScanRows * kScanRows;
}

struct ConsLayout {
  __nv_bfloat16* X;        // bf16 image of the caller's rows
  float* xnorm;
  __nv_bfloat16* Y;        // kept rows so far, compacted, followed by the current band
  float* ynorm;
  uint32_t* mask;          // bit matrix of the current band: row = Y row - 512 (K / 512), columns in Y row numbers
  int64_t words_per_row;
  unsigned long long* kept[2];   // block-local tagged kept-words of the current / next band
  int kept_words;
  uint2* unc;
  int32_t unc_cap;
  int32_t* counters;       // [0] uncertain count, [2] inexact, [4..5] dyn K (two parities), [8..11] running stats
  size_t bytes;
};

// default capacity of the near-threshold pair list (drained after every band)
static int64_t cons_default_cap(int64_t n, int band) {
  return 64 * ((n < (int64_t)band ? n : (int64_t)band) + 1024) + (1 << 20);
}

static ConsLayout cons_layout(void* ws, size_t ws_bytes, int64_t n, int d, int band, int64_t unc_cap = 0) {
  Carver c(ws, ws_bytes);
  ConsLayout L{};
  L.X = c.take<__nv_bfloat16>((size_t)n * d);
  L.xnorm = c.take<float>((size_t)n);
  L.Y = c.take<__nv_bfloat16>((size_t)n * d);
  L.ynorm = c.take<float>((size_t)n);
  L.words_per_row = (n + kTcBN - 1) / kTcBN * (kTcBN / 32);
  // the current band only: rows [512 (K / 512), K + band) rounded up to the 256-row blocks the contraction writes
  const int64_t mask_rows = (n < (int64_t)band ? n : (int64_t)band) + 1024;
  L.mask = c.take<uint32_t>((size_t)mask_rows * L.words_per_row);
  L.kept_words = (band / kScanRows + 1) * kScanWords;
  L.kept[0] = c.take<unsigned long long>((size_t)L.kept_words);
  L.kept[1] = c.take<unsigned long long>((size_t)L.kept_words);
  // the list is drained after every band
  int64_t cap = unc_cap > 0 ? unc_cap : cons_default_cap(n, band);
  if (cap > 0x7fffff00ll) cap = 0x7fffff00ll;
  L.unc_cap = (int32_t)cap;
  L.unc = c.take<uint2>((size_t)cap);
  L.counters = c.take<int32_t>(64);
  L.bytes = c.used();
  return L;
}

}  // namespace hippo

extern "C" {

size_t hippo_consolidate_workspace_bytes(int64_t n, int32_t d) {
  return hippo_consolidate_ex_workspace_bytes(n, d, 0, 0);
}

hippo_status hippo_consolidate(const float* feats, int64_t n, int32_t d, float gamma, float band_exact,
                               float band_inexact, int64_t* out_keep, int32_t* out_count,
                               int32_t* out_stats, void* ws, size_t ws_bytes, void* stream) {
  return hippo_consolidate_ex(feats, n, d, gamma, band_exact, band_inexact, 0, 0, out_keep, out_count, out_stats, ws,
                              ws_bytes, stream);
}

size_t hippo_consolidate_ex_workspace_bytes(int64_t n, int32_t d, int32_t band_rows, int64_t uncertain_cap) {
  if (n <= 2 || d <= 0) return 256;
  return hippo::cons_layout(nullptr, 0, n, d, hippo::cons_band(band_rows), uncertain_cap).bytes;
}

hippo_status hippo_consolidate_ex(const float* feats, int64_t n, int32_t d, float gamma, float band_exact,
                                  float band_inexact, int32_t band_rows, int64_t uncertain_cap, int64_t* out_keep,
                                  int32_t* out_count, int32_t* out_stats, void* ws, size_t ws_bytes, void* stream) {
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
  HIPPO_REQUIRE(band_rows >= 0 && uncertain_cap >= 0, "hippo_consolidate: negative band_rows / uncertain_cap");
  const int band = cons_band(band_rows);
  ConsLayout L = cons_layout(ws, ws_bytes, n, d, band, uncertain_cap);
  if (ws == nullptr || ((uintptr_t)ws & 255) || L.bytes > ws_bytes) {
    set_error("hippo_consolidate: workspace of %zu bytes needed (256-byte aligned), got %zu", L.bytes, ws_bytes);
    return HIPPO_E_WORKSPACE;
  }
  HIPPO_CUDA(cudaMemsetAsync(L.counters, 0, 64 * sizeof(int32_t), s));
  st = hippo_bank_build(feats, HIPPO_F32, n, d, d, L.X, L.xnorm, L.counters + 2, stream);
  if (st != HIPPO_OK) return st;
  int32_t* dyn = L.counters + 4;
  int32_t* stats = L.counters + 8;

  TcMaskArgs a{};
  a.feats_bf16 = L.Y;
  a.norm = L.ynorm;
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

  const int adv_grid = sm_count() * 8;
  const int scan_grid = band / kScanRows + 1;
  const bool dbg_on = getenv("HIPPO_SCAN_DEBUG") != nullptr;
  const bool timing = getenv("HIPPO_CONS_TIMING") != nullptr;
  struct Stamp { cudaEvent_t e; int stage; };
  std::vector<Stamp> stamps;
  auto stamp = [&](int stage) {           // stage = what ran SINCE the previous stamp
    if (!timing) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    stamps.push_back({e, stage});
  };
  stamp(3);
  int par = 1;                      // dyn[par] = K before the band being finished
  int64_t prev_r0 = 0;
  int prev_rows = 0;
  int iband = 0;
  for (int64_t r0 = 0;; r0 += band, ++iband) {
    const int rows = (int)(r0 < n ? (n - r0 < band ? n - r0 : band) : 0);
    // kept-words: band i publishes into kept[i & 1]; the advance after band i-1 reads kept[(i-1) & 1] and clears kept[i & 1]
    cons_advance_kernel<<<adv_grid, kAdvThreads, 0, s>>>(L.X, L.xnorm, d, L.Y, L.ynorm, out_keep,
                                                         L.kept[(iband + 1) & 1], L.kept[iband & 1], L.kept_words, dyn,
                                                         par, prev_r0, prev_rows, r0, rows, L.counters + 0, L.unc_cap,
                                                         stats);
    HIPPO_CUDA(cudaGetLastError());
    stamp(3);
    par ^= 1;                       // dyn[par] = K before this band
    if (rows == 0) break;
    a.dyn_k = dyn + par;
    a.band_rows = rows;
    st = tc_mask_launch(a, s);
    if (st != HIPPO_OK) return st;
    stamp(0);
    recheck_kernel<<<sm_count() * 4, 256, 0, s>>>(feats, L.xnorm, d, out_keep, gamma, L.unc, L.counters + 0,
                                                   L.unc_cap, L.mask, L.words_per_row, dyn + par);
    HIPPO_CUDA(cudaGetLastError());
    stamp(1);
    unsigned long long* dbg = nullptr;
    if (dbg_on) { cudaMalloc(&dbg, (size_t)scan_grid * 64); cudaMemset(dbg, 0, (size_t)scan_grid * 64); }
    greedy_scan_kernel<<<scan_grid, kScanThreads, 0, s>>>(L.mask, L.words_per_row, dyn + par, rows, L.kept[iband & 1], dbg);
    HIPPO_CUDA(cudaGetLastError());
    stamp(2);
    if (dbg) {
      cudaStreamSynchronize(s);
      std::vector<unsigned long long> h((size_t)scan_grid * 8);
      cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost);
      cudaFree(dbg);
      const unsigned long long t0 = h[0];
      for (int i = 0; i < scan_grid; ++i)
        if (iband % 8 == 0 && h[i * 8 + 3])
          fprintf(stderr, "[scan] band %d blk %2d start %7.2f bulk_done %7.2f phaseA %7.2f (%llu rounds) preds %7.2f published %7.2f us (rounds %llu)\n",
                  iband, i, (h[i * 8] - t0) / 1e3, (h[i * 8 + 1] - t0) / 1e3, (h[i * 8 + 5] - t0) / 1e3, h[i * 8 + 6],
                  (h[i * 8 + 2] - t0) / 1e3, (h[i * 8 + 3] - t0) / 1e3, h[i * 8 + 4]);
    }
    prev_r0 = r0;
    prev_rows = rows;
  }
  cons_finish_kernel<<<1, 1, 0, s>>>(dyn, par, L.counters + 2, stats, out_count, out_stats);
  HIPPO_CUDA(cudaGetLastError());
  if (timing) {
    stamp(3);
    cudaStreamSynchronize(s);
    for (double& v : g_cons_ms) v = 0.0;
    for (size_t i = 1; i < stamps.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, stamps[i - 1].e, stamps[i].e);
      g_cons_ms[stamps[i].stage] += ms;
    }
    for (auto& x : stamps) cudaEventDestroy(x.e);
  }
  return HIPPO_OK;
}

void hippo_debug_consolidate_timing(double* out4_host) {
  for (int i = 0; i < 4; ++i) out4_host[i] = hippo::g_cons_ms[i];
}

}  // extern "C"
