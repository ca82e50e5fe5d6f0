// hippo_consolidate: greedy cosine-redundancy filter (reference: _select_key_frames, hm:944-967).
//
// The reference forms the whole N x N similarity matrix (hm:952) and then walks the rows: row i is kept iff
// every KEPT earlier row j has sim(i, j) < gamma (hm:958-961).  Similarities against rows that were dropped
// are never looked at -- and on video-like input most rows are dropped.  So the rows are processed in BANDS
// of kConsBand rows, and a band is only contracted against
//     (R) the rows kept so far, held compacted at the front of a second bf16 matrix Y  (a rectangle), and
//     (T) itself                                                                        (a lower triangle),
// which is exactly the set of pairs the greedy rule can consult.  With K rows kept out of N the tensor work
// drops from N^2/2 to about N K / 2 + N band / 2 pairs (10x at 6% kept); when everything is kept it is the
// full triangle again.  The decisions are identical either way.
//
// Only R depends on earlier decisions (it needs the kept rows of the band before), T does not -- so the
// triangle of band b + 1 runs UNDER the decision chain of band b.  Two streams:
//   tensor stream (internal)      T(0) rT(0) | R(0) T(1) rT(1) | R(1) T(2) rT(2) | ...
//   caller's stream               bank build | rR(0) scan(0) compact(0) | rR(1) scan(1) compact(1) | ...
//     T(b)        sim_tc EPI_MASK triangle of band b, rows straight from X (bf16 image of the caller's rows) ->
//                 band-local bit matrix Mt[b & 1]; on 65 of the 74 CTA pairs, so that the chain's kernels find SMs
//     R(b)        sim_tc EPI_MASK band b x Y[0, K) -> one conflict flag per band row             (hm:952, hm:960)
//     rT / rR     pairs within `band` of gamma re-evaluated from the fp32 rows
//     scan(b)     row i kept iff no kept row before it has its bit set: the flag of R(b) (all of Y is kept), then the
//                 sequential part over Mt                                                        (hm:958-961)
//     compact(b)  the band's kept rows are appended to Y, K grows, out_keep receives their row numbers
// R(b + 1) waits for compact(b); everything else of band b + 1 is done by then.  All ordering is by stream events,
// the extent of every launch is read from device memory (K), nothing synchronises with the host.
//
// The N x N fp32 matrix of the reference (40 GB at N = 100k) is never formed: the triangle's bits live in two
// band-local matrices (2 x 8 MB), and of the rectangle only the OR over a band row's bits is kept (one flag per
// row: every column there is a kept row, so a single conflict settles the row).
#include "common.cuh"
#include "sim_tc.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace hippo {

constexpr int kConsBandDefault = 8192;

// ---- exact re-evaluation of near-threshold pairs ---------------------------------
// One warp per pair (i, j): i = row of the band (original row r0 + i), j = column -- a row of the same band
// (triangle: original row r0 + j) or a row of Y (rectangle: original row yidx[j]).  The rows are taken from the
// caller's fp32 matrix, normalised element-wise in fp32 exactly like hm:951 (x / |x| with the fp32 norm), the
// products are accumulated in fp64, and the bit becomes !(sim < gamma) -- the best available stand-in for the
// reference's fp32 sgemm value.
__global__ void __launch_bounds__(256) recheck_kernel(const float* __restrict__ feats,
                                                      const float* __restrict__ norm, int d, int64_t r0,
                                                      const int64_t* __restrict__ yidx /* null: triangle */,
                                                      float gamma, const uint2* __restrict__ pairs,
                                                      const int32_t* __restrict__ count, int32_t cap,
                                                      uint32_t* __restrict__ mask, int64_t words_per_row,
                                                      int32_t* __restrict__ rowhit /* rectangle */,
                                                      int32_t* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int32_t np = *count;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&stats[0], np < cap ? np : cap);
    if (np > cap) atomicOr(&stats[1], 1);
  }
  if (np > cap) np = cap;
  for (int64_t pi = warp; pi < np; pi += nwarps) {
    const uint2 pr = pairs[pi];
    const int64_t ra = r0 + pr.x, rb = yidx ? yidx[pr.y] : r0 + pr.y;
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
      if (rowhit != nullptr) {                       // rectangle: the column is a kept row, a conflict drops the band row
        if (bit) rowhit[pr.x] = 1;
      } else {                                       // triangle: the contraction left the bit clear
        if (bit) atomicOr(mask + (int64_t)pr.x * words_per_row + (pr.y >> 5), 1u << (pr.y & 31));
      }
    }
  }
}

// ---- greedy scan over the bit matrices of one band --------------------------------------
// A row of the band is dropped at once if any bit of its row of the RECTANGLE matrix is set (every row of Y is a kept
// row, final).  The rest is the sequential part, over the band-local TRIANGLE matrix: CTA x owns band rows
// [512 x, 512 x + 512), one thread per row.  The chain over blocks is sequential (a row's fate depends on which
// earlier rows were KEPT), so the kernel is bound by the hand-off latency between consecutive blocks; everything
// that does not depend on the predecessors' results is done ahead:
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

__global__ void __launch_bounds__(kScanThreads, 1) greedy_scan_kernel(const uint32_t* __restrict__ mask /*triangle*/,
                                                                      int64_t words_per_row,
                                                                      const int32_t* __restrict__ rowhit,
                                                                      const int32_t* __restrict__ dyn_k, int band_rows,
                                                                      unsigned long long* kept /*[grid * 16], zeroed*/,
                                                                      unsigned long long* dbg) {
  __shared__ __align__(16) uint32_t s_kept[2][kScanWords];
  __shared__ __align__(16) uint32_t s_undec[2][kScanWords];
  __shared__ __align__(16) uint32_t s_kw[2][kScanWords];            // bulk ring
  __shared__ __align__(16) uint32_t s_pk[kScanPreds][kScanWords];   // predecessors' kept-words
  (void)dyn_k;
  const int64_t n = band_rows;
  constexpr int b0 = 0;                            // blocks are numbered inside the band
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t row = (int64_t)blockIdx.x * kScanRows + tid;   // row of the band
  if ((int64_t)b * kScanRows >= n) return;         // nobody waits for a block past the end
  const bool live = row < n;
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
  uint32_t hit = 0;
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
    // a conflict with a row kept before this band (all of them final) drops the row: the rectangle kernel and the
    // re-evaluation of its near-threshold pairs left one flag per band row
    hit = (uint32_t)rowhit[row];
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

// ---- compaction of a band's kept rows -----------------------------------------
// dyn[par_in] = K before the band, whose rows are X[r0, r0 + rows) and whose kept-words are kept_band (bit = row of
// the band).  The kept rows are appended to Y[K, K'), out_keep[K ..) receives their original row numbers
// (ascending: hm:967), K' goes to dyn[par_in ^ 1], and the kept-words of the band after next are cleared.
// Every CTA recomputes the (short) prefix over the kept-words on its own.
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
constexpr int kAdvMaxWords = 1024;   // band of at most 32k rows

__global__ void __launch_bounds__(kAdvThreads) cons_compact_kernel(
    const __nv_bfloat16* __restrict__ X, const float* __restrict__ xnorm, int d, __nv_bfloat16* __restrict__ Y,
    float* __restrict__ ynorm, int64_t* __restrict__ yidx, const unsigned long long* __restrict__ kept_band,
    unsigned long long* __restrict__ kept_clear, int kept_words, int32_t* dyn, int par_in, int64_t r0, int rows) {
  __shared__ int s_pref[kAdvMaxWords + 1];
  __shared__ uint32_t s_bits[kAdvMaxWords];
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = dyn[par_in];
  const int nwords = (rows + 31) / 32;
  for (int w = tid; w < nwords; w += kAdvThreads) {
    uint32_t bits = (uint32_t)kept_band[w];
    const int end = rows - w * 32;                                     // rows past the band
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
  const int K2 = K + s_pref[nwords];

  const int64_t warp = (int64_t)blockIdx.x * (kAdvThreads / 32) + (tid >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kAdvThreads / 32);
  const int vec_per_row = d / 8;                                       // 16-byte vectors per bf16 row
  for (int64_t t = warp; t < rows; t += nwarps) {                      // one warp per kept row
    const uint32_t bits = s_bits[t >> 5];
    if (!((bits >> (t & 31)) & 1u)) continue;
    const int rank = s_pref[t >> 5] + __popc(bits & ((1u << (t & 31)) - 1u));
    const int64_t src = r0 + t, dst = (int64_t)K + rank;
    copy_row(reinterpret_cast<const uint4*>(X + src * d), reinterpret_cast<uint4*>(Y + dst * d), vec_per_row, lane);
    if (lane == 0) { ynorm[dst] = xnorm[src]; yidx[dst] = src; }
  }
  for (int64_t i = (int64_t)blockIdx.x * kAdvThreads + tid; i < kept_words; i += (int64_t)gridDim.x * kAdvThreads)
    kept_clear[i] = 0;
  if (blockIdx.x == 0 && tid == 0) dyn[par_in ^ 1] = K2;
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

// HIPPO_CONS_TIMING: CUDA-event time per stage of the last hippo_consolidate call on this thread
// {T + R launches on the tensor stream, fp32 re-evaluation, greedy scan, bank build + compaction}; with the stages
// overlapped the four do not add up to the call's duration any more.  Read with hippo_debug_consolidate_timing.
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
  __nv_bfloat16* Y;        // kept rows so far, compacted
  float* ynorm;
  uint32_t* mt[2];         // triangle bit matrix of a band (band-local rows and columns), two bands in flight
  int64_t mt_words;
  int32_t* rowhit;         // rectangle: one flag per band row (conflict with a row kept before the band)
  unsigned long long* kept[2];   // tagged kept-words of the current / next band
  int kept_words;
  uint2* unc_t;            // near-threshold pairs of the triangle / of the rectangle
  uint2* unc_r;
  int32_t unc_cap;
  int32_t* counters;       // [0] triangle list count, [1] rectangle list count, [2] inexact, [4..5] dyn K (two parities), [8..11] statistics
  size_t bytes;
};

// default capacity of the near-threshold pair lists (each drained after every band)
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
  const int64_t brows = n < (int64_t)band ? n : (int64_t)band;
  const int64_t brows_pad = (brows + kTcBN - 1) / kTcBN * kTcBN;          // the contraction writes whole 256-row blocks' rows < na only, pad anyway
  L.mt_words = (brows + kScanRows - 1) / kScanRows * kScanWords;         // whole 512-column scan blocks
  L.mt[0] = c.take<uint32_t>((size_t)brows_pad * L.mt_words);
  L.mt[1] = c.take<uint32_t>((size_t)brows_pad * L.mt_words);
  L.rowhit = c.take<int32_t>((size_t)brows_pad);
  L.kept_words = (band / kScanRows + 1) * kScanWords;
  L.kept[0] = c.take<unsigned long long>((size_t)L.kept_words);
  L.kept[1] = c.take<unsigned long long>((size_t)L.kept_words);
  int64_t cap = unc_cap > 0 ? unc_cap : cons_default_cap(n, band);
  if (cap > 0x7fffff00ll) cap = 0x7fffff00ll;
  L.unc_cap = (int32_t)cap;
  L.unc_t = c.take<uint2>((size_t)cap);
  L.unc_r = c.take<uint2>((size_t)cap);
  L.counters = c.take<int32_t>(64);
  L.bytes = c.used();
  return L;
}

// the internal tensor stream and the events that order it against the caller's stream: created once per host thread
static cudaStream_t cons_side_stream() {
  static thread_local cudaStream_t st = nullptr;
  if (st == nullptr && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) st = nullptr;
  return st;
}
static cudaEvent_t cons_event(size_t i) {
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
  cudaStream_t ts = cons_side_stream();
  HIPPO_REQUIRE(ts != nullptr, "hippo_consolidate: could not create the internal stream");
  // HIPPO_CONS_OVERLAP=0: everything on the caller's stream, in dependency order (debugging aid)
  const bool overlap = !(getenv("HIPPO_CONS_OVERLAP") && atoi(getenv("HIPPO_CONS_OVERLAP")) == 0);
  if (!overlap) ts = s;
  size_t nev = 0;
  auto next_event = [&]() { return cons_event(nev++); };

  const bool timing = getenv("HIPPO_CONS_TIMING") != nullptr;
  struct Span { cudaEvent_t a, b; int stage; };
  std::vector<Span> spans;
  auto span_begin = [&](cudaStream_t q, int stage) -> int {
    if (!timing) return -1;
    Span sp{};
    cudaEventCreate(&sp.a);
    cudaEventCreate(&sp.b);
    sp.stage = stage;
    cudaEventRecord(sp.a, q);
    spans.push_back(sp);
    return (int)spans.size() - 1;
  };
  auto span_end = [&](cudaStream_t q, int id) { if (id >= 0) cudaEventRecord(spans[id].b, q); };

  int sp = span_begin(s, 3);
  HIPPO_CUDA(cudaMemsetAsync(L.counters, 0, 64 * sizeof(int32_t), s));
  HIPPO_CUDA(cudaMemsetAsync(L.kept[0], 0, (size_t)L.kept_words * 8, s));
  HIPPO_CUDA(cudaMemsetAsync(L.kept[1], 0, (size_t)L.kept_words * 8, s));
  HIPPO_CUDA(cudaMemsetAsync(L.rowhit, 0, (size_t)(n < band ? n : band) * sizeof(int32_t), s));   // band 0 has no rectangle
  st = hippo_bank_build(feats, HIPPO_F32, n, d, d, L.X, L.xnorm, L.counters + 2, stream);
  if (st != HIPPO_OK) return st;
  span_end(s, sp);
  int32_t* dyn = L.counters + 4;
  int32_t* stats = L.counters + 8;
  cudaEvent_t e0 = next_event();
  HIPPO_REQUIRE(e0 != nullptr, "hippo_consolidate: could not create events");
  if (overlap) {
    HIPPO_CUDA(cudaEventRecord(e0, s));
    HIPPO_CUDA(cudaStreamWaitEvent(ts, e0, 0));
  }

  TcMaskArgs a{};
  a.a_rows = L.X;
  a.a_total = n;
  a.d = d;
  a.gamma = gamma;
  a.band_exact = band_exact;
  a.band_inexact = band_inexact;
  a.inexact = L.counters + 2;
  a.uncertain_cap = L.unc_cap;

  const int compact_grid = sm_count() * 2;
  // the rectangle's re-evaluation runs on the few SMs the next triangle leaves free: a grid that fits them in one wave
  // (fp32 rows, ~16k pairs per band: 2.36 -> 2.32 ms; 32 - 296 CTAs make no difference on bf16-exact rows, ~125 pairs)
  const int recheck_grid_r = 64;
  const int scan_grid = band / kScanRows + 1;
  const int nbands = (int)((n + band - 1) / band);
  const bool dbg_on = getenv("HIPPO_SCAN_DEBUG") != nullptr;
  // the triangle kernels run beside the chain's kernels (scan: up to 17 CTAs of 512 threads; re-evaluation and
  // compaction: many short CTAs), which do not fit on an SM next to a tcgen05 CTA: leave them a few SMs
  const int pairs_all = sm_count() / 2;
  int pairs_t = 0;
  if (overlap && nbands > 1 && pairs_all > 12) {
    // the fewest pairs that still finish the triangle's tiles in as few rounds as the whole machine would need
    // (528 tiles of an 8,192-row band: 8 rounds on 74 pairs -- and on 66); if that leaves the chain fewer than four
    // pairs' SMs, one round more
    const int hb = (int)(((n < band ? n : band) + kTcBN - 1) / kTcBN);
    const int tiles = hb * (hb + 1) / 2;
    int rounds = (tiles + pairs_all - 1) / pairs_all;
    pairs_t = (tiles + rounds - 1) / rounds;
    if (pairs_t > pairs_all - 4) { ++rounds; pairs_t = (tiles + rounds - 1) / rounds; }
    if (pairs_t < 1) pairs_t = 1;
  }

  // T(b) + its re-evaluation on the tensor stream; the bit matrix Mt[b & 1] was last read by scan(b - 2)
  std::vector<cudaEvent_t> e_scan(nbands, nullptr);
  auto launch_triangle = [&](int b) -> hippo_status {
    const int64_t r0 = (int64_t)b * band;
    const int rows = (int)(n - r0 < band ? n - r0 : band);
    if (overlap && b >= 2) HIPPO_CUDA(cudaStreamWaitEvent(ts, e_scan[b - 2], 0));
    int id = span_begin(ts, 0);
    HIPPO_CUDA(cudaMemsetAsync(L.counters + 0, 0, sizeof(int32_t), ts));
    TcMaskArgs t = a;
    t.rect = false;
    t.b_rows = L.X;
    t.b_total = n;
    t.a_row0 = t.b_row0 = r0;
    t.na = rows;
    t.anorm = t.bnorm = L.xnorm + r0;
    t.dyn_k = nullptr;
    t.mask = L.mt[b & 1];
    t.words_per_row = L.mt_words;
    t.uncertain = L.unc_t;
    t.uncertain_count = L.counters + 0;
    t.max_pairs = pairs_t;
    hippo_status r = tc_mask_launch(t, ts);
    if (r != HIPPO_OK) return r;
    span_end(ts, id);
    id = span_begin(ts, 1);
    recheck_kernel<<<sm_count() * 2, 256, 0, ts>>>(feats, L.xnorm, d, r0, nullptr, gamma, L.unc_t, L.counters + 0, L.unc_cap,
                                                   L.mt[b & 1], L.mt_words, nullptr, stats);
    HIPPO_CUDA(cudaGetLastError());
    span_end(ts, id);
    return HIPPO_OK;
  };

  st = launch_triangle(0);
  if (st != HIPPO_OK) return st;
  int par = 0;                      // dyn[par] = K before the band being processed (0 for the first)
  for (int b = 0; b < nbands; ++b) {
    const int64_t r0 = (int64_t)b * band;
    const int rows = (int)(n - r0 < band ? n - r0 : band);
    // ---- tensor stream: R(b) (needs compact(b - 1): the caller's stream is joined first), then T(b + 1) ----
    if (b > 0) {
      if (overlap) {
        cudaEvent_t ec = next_event();
        HIPPO_REQUIRE(ec != nullptr, "hippo_consolidate: could not create events");
        HIPPO_CUDA(cudaEventRecord(ec, s));
        HIPPO_CUDA(cudaStreamWaitEvent(ts, ec, 0));
      }
      int id = span_begin(ts, 0);
      HIPPO_CUDA(cudaMemsetAsync(L.counters + 1, 0, sizeof(int32_t), ts));
      HIPPO_CUDA(cudaMemsetAsync(L.rowhit, 0, (size_t)rows * sizeof(int32_t), ts));
      TcMaskArgs r = a;
      r.rect = true;
      r.b_rows = L.Y;
      r.b_total = n;
      r.a_row0 = r0;
      r.b_row0 = 0;
      r.na = rows;
      r.anorm = L.xnorm + r0;
      r.bnorm = L.ynorm;
      r.dyn_k = dyn + par;
      r.mask = nullptr;
      r.words_per_row = 0;
      r.rowhit = L.rowhit;
      r.uncertain = L.unc_r;
      r.uncertain_count = L.counters + 1;
      r.max_pairs = 0;
      st = tc_mask_launch(r, ts);
      if (st != HIPPO_OK) return st;
      span_end(ts, id);
    }
    if (overlap) {
      cudaEvent_t er = next_event();
      HIPPO_REQUIRE(er != nullptr, "hippo_consolidate: could not create events");
      HIPPO_CUDA(cudaEventRecord(er, ts));          // T(b), its re-evaluation and R(b) are done
      HIPPO_CUDA(cudaStreamWaitEvent(s, er, 0));
    }
    if (b + 1 < nbands) {
      st = launch_triangle(b + 1);
      if (st != HIPPO_OK) return st;
    }
    // ---- caller's stream: re-evaluation of R(b), scan(b), compact(b) ----
    if (b > 0) {
      int id = span_begin(s, 1);
      recheck_kernel<<<recheck_grid_r, 256, 0, s>>>(feats, L.xnorm, d, r0, out_keep, gamma, L.unc_r, L.counters + 1, L.unc_cap,
                                                    nullptr, 0, L.rowhit, stats);
      HIPPO_CUDA(cudaGetLastError());
      span_end(s, id);
    }
    unsigned long long* dbg = nullptr;
    if (dbg_on) { cudaMalloc(&dbg, (size_t)scan_grid * 64); cudaMemset(dbg, 0, (size_t)scan_grid * 64); }
    int id = span_begin(s, 2);
    greedy_scan_kernel<<<scan_grid, kScanThreads, 0, s>>>(L.mt[b & 1], L.mt_words, L.rowhit, dyn + par, rows,
                                                          L.kept[b & 1], dbg);
    HIPPO_CUDA(cudaGetLastError());
    span_end(s, id);
    if (overlap) {
      e_scan[b] = next_event();
      HIPPO_REQUIRE(e_scan[b] != nullptr, "hippo_consolidate: could not create events");
      HIPPO_CUDA(cudaEventRecord(e_scan[b], s));
    }
    if (dbg) {
      cudaStreamSynchronize(s);
      std::vector<unsigned long long> h((size_t)scan_grid * 8);
      cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost);
      cudaFree(dbg);
      const unsigned long long t0 = h[0];
      for (int i = 0; i < scan_grid; ++i)
        if (b % 8 == 0 && h[i * 8 + 3])
          fprintf(stderr, "[scan] band %d blk %2d start %7.2f bulk_done %7.2f phaseA %7.2f (%llu rounds) preds %7.2f published %7.2f us (rounds %llu)\n",
                  b, i, (h[i * 8] - t0) / 1e3, (h[i * 8 + 1] - t0) / 1e3, (h[i * 8 + 5] - t0) / 1e3, h[i * 8 + 6],
                  (h[i * 8 + 2] - t0) / 1e3, (h[i * 8 + 3] - t0) / 1e3, h[i * 8 + 4]);
    }
    id = span_begin(s, 3);
    // kept-words: band b publishes into kept[b & 1]; this launch reads them and clears the OTHER buffer for band
    // b + 1 (last read by compact(b - 1), which is done)
    cons_compact_kernel<<<compact_grid, kAdvThreads, 0, s>>>(L.X, L.xnorm, d, L.Y, L.ynorm, out_keep, L.kept[b & 1],
                                                             L.kept[(b + 1) & 1], L.kept_words, dyn, par, r0, rows);
    HIPPO_CUDA(cudaGetLastError());
    span_end(s, id);
    par ^= 1;
  }
  cons_finish_kernel<<<1, 1, 0, s>>>(dyn, par, L.counters + 2, stats, out_count, out_stats);
  HIPPO_CUDA(cudaGetLastError());
  if (timing) {
    cudaStreamSynchronize(s);
    cudaStreamSynchronize(ts);
    for (double& v : g_cons_ms) v = 0.0;
    for (auto& x : spans) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, x.a, x.b);
      g_cons_ms[x.stage] += ms;
      cudaEventDestroy(x.a);
      cudaEventDestroy(x.b);
    }
  }
  return HIPPO_OK;
}

void hippo_debug_consolidate_timing(double* out4_host) {
  for (int i = 0; i < 4; ++i) out4_host[i] = hippo::g_cons_ms[i];
}

}  // extern "C"
