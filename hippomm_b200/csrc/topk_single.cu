// hippo_topk_single / hippo_topk_merge: one-query cosine top-k over the device bank.
//
// Reference: top_k_cosine_similarity (vo:151-188): sims = b.a / (|b| |a|), argsort, last k
// reversed.  Here: a bandwidth-bound GEMV.  Each warp streams groups of 4 bank rows with
// 128-bit non-allocating loads (the bank is read exactly once), multiplies against the fp32
// query staged in shared memory, reduces with shuffles, and keeps a warp-private top-k list
// that only rows passing a threshold test ever touch.  Per-block lists are merged by a
// selection pass; a second tiny kernel merges the per-block results.
#include "exchange.cuh"
#include <cstdlib>

namespace hippo {

constexpr int kSingleThreads = 512;
constexpr int kSingleWarps = kSingleThreads / 32;
constexpr int kRowsPerIter = 4;

// r-th best = largest key strictly below the previous winner; keys are unique per row.
// cand[i * stride] for i in [0, ncand); returns 0 when nothing is left.
__device__ __forceinline__ uint64_t warp_next_best(const uint64_t* cand, int ncand, uint64_t below,
                                                   int lane) {
  uint64_t best = 0;
  for (int i = lane; i < ncand; i += 32) {
    uint64_t c = cand[i];
    if (c < below && c > best) best = c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  return best;
}

// Filter threshold for t = dot * (1/|b|) given the current k-th best score:
// anything that could reach `kth` after the exact IEEE evaluation must pass !(t < thr).
__device__ __forceinline__ float filter_threshold(uint64_t kth_key, float an) {
  if (kth_key == 0) return -INFINITY;                       // list not full yet
  uint32_t ord = (uint32_t)(kth_key >> 32);
  if (ord == 0xffffffffu) return INFINITY;                  // k NaNs already: only NaN/inf pass
  float lo = ord_to_score(ord) * an;
  return lo - fabsf(lo) * 9.5367431640625e-07f - 1e-37f;    // 2^-20 relative slack
}

// Tail of the sharded single-query search, run by the LAST CTA of the GEMV to finish (ticket counter): merge of the
// per-block lists, push of the k keys into every peer's gather buffer over NVLink, flag exchange, merge of the
// world x k candidates -- the whole query is ONE launch per rank (GEMV, merge kernel and exchange kernel were three:
// ~0.16 ms of a 0.36 ms shard pass at 8 GPUs).  world == 1 degenerates to the local merge.
struct SingleTail {
  uint32_t* ticket;                  // zero before the launch; the last CTA resets it
  unsigned char* const* peer_bases;  // device array [world] (see hippo_topk_exchange_merge)
  size_t slot_stride;
  int rank, world;
  uint32_t epoch;
  int64_t* out_idx; float* out_score; uint64_t* out_key;
};

template <int CH, bool kFused>  // CH = d/256 when d is a multiple of 256 and <= 4*256; 0 = generic
__global__ void __launch_bounds__(kSingleThreads, 2)
topk_single_kernel(const __nv_bfloat16* __restrict__ bank, const float* __restrict__ norm, int64_t n,
                   int d, const float* __restrict__ q, int k, int64_t row_base,
                   const uint64_t* __restrict__ after_key, uint64_t* __restrict__ part /*[grid][k]*/,
                   const SingleTail tail) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // query, permuted so that a lane's two float4 loads are conflict free:
  // element e = c*256 + lane*8 + h*4 + j  ->  sq[((c*2 + h)*32 + lane)*4 + j]
  float* sq = reinterpret_cast<float*>(smem_raw);
  const int dpad = (d + 255) / 256 * 256;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + (size_t)dpad * 4);  // [warps][k]
  __shared__ float s_an;

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  // stage query, compute |a| the way the reference does (fp32 sqrt of the sum of squares)
  double ss = 0.0;
  for (int e = threadIdx.x; e < dpad; e += blockDim.x) {
    float v = e < d ? q[e] : 0.f;
    ss += (double)v * (double)v;
    int c = e >> 8, l = (e >> 3) & 31, h = (e >> 2) & 1, j = e & 3;
    sq[(((c * 2 + h) * 32) + l) * 4 + j] = v;
  }
  ss = warp_sum(ss);
  __shared__ double s_part[kSingleWarps];
  if (lane == 0) s_part[wid] = ss;
  for (int i = lane; i < k; i += 32) lists[wid * k + i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kSingleWarps; ++i) t += s_part[i];
    s_an = __fsqrt_rn((float)t);
  }
  __syncthreads();
  const float an = s_an;
  const uint64_t below = after_key ? after_key[0] : ~0ull;
  uint64_t* mylist = lists + wid * k;
  float thr = -INFINITY;

  const int chunks = dpad >> 8;
  const int64_t ngroups = (n + kRowsPerIter - 1) / kRowsPerIter;
  const int64_t gwarp = (int64_t)blockIdx.x * kSingleWarps + wid;
  const int64_t gstride = (int64_t)gridDim.x * kSingleWarps;

  for (int64_t g = gwarp; g < ngroups; g += gstride) {
    const int64_t r0 = g * kRowsPerIter;
    float acc[kRowsPerIter];
#pragma unroll
    for (int r = 0; r < kRowsPerIter; ++r) acc[r] = 0.f;
    const __nv_bfloat16* rp[kRowsPerIter];
#pragma unroll
    for (int r = 0; r < kRowsPerIter; ++r) {
      int64_t row = r0 + r < n ? r0 + r : n - 1;
      rp[r] = bank + row * (int64_t)d + lane * 8;
    }
    if constexpr (CH > 0) {
      uint4 w[CH][kRowsPerIter];
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int r = 0; r < kRowsPerIter; ++r) w[c][r] = ldg_stream(rp[r] + c * 256);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float4 qa = *reinterpret_cast<const float4*>(&sq[((c * 2 + 0) * 32 + lane) * 4]);
        const float4 qb = *reinterpret_cast<const float4*>(&sq[((c * 2 + 1) * 32 + lane) * 4]);
#pragma unroll
        for (int r = 0; r < kRowsPerIter; ++r) {
          const uint4 x = w[c][r];
          float a = acc[r];
          a = fmaf(bf16lo(x.x), qa.x, a); a = fmaf(bf16hi(x.x), qa.y, a);
          a = fmaf(bf16lo(x.y), qa.z, a); a = fmaf(bf16hi(x.y), qa.w, a);
          a = fmaf(bf16lo(x.z), qb.x, a); a = fmaf(bf16hi(x.z), qb.y, a);
          a = fmaf(bf16lo(x.w), qb.z, a); a = fmaf(bf16hi(x.w), qb.w, a);
          acc[r] = a;
        }
      }
    } else {
      for (int c = 0; c < chunks; ++c) {
        if (c * 256 + lane * 8 < d) {
          const float4 qa = *reinterpret_cast<const float4*>(&sq[((c * 2 + 0) * 32 + lane) * 4]);
          const float4 qb = *reinterpret_cast<const float4*>(&sq[((c * 2 + 1) * 32 + lane) * 4]);
          uint4 w[kRowsPerIter];
#pragma unroll
          for (int r = 0; r < kRowsPerIter; ++r) w[r] = ldg_stream(rp[r] + c * 256);
#pragma unroll
          for (int r = 0; r < kRowsPerIter; ++r) {
            const uint4 x = w[r];
            float a = acc[r];
            a = fmaf(bf16lo(x.x), qa.x, a); a = fmaf(bf16hi(x.x), qa.y, a);
            a = fmaf(bf16lo(x.y), qa.z, a); a = fmaf(bf16hi(x.y), qa.w, a);
            a = fmaf(bf16lo(x.z), qb.x, a); a = fmaf(bf16hi(x.z), qb.y, a);
            a = fmaf(bf16lo(x.w), qb.z, a); a = fmaf(bf16hi(x.w), qb.w, a);
            acc[r] = a;
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerIter; ++r) acc[r] = warp_sum(acc[r]);

    // lane r owns row r0 + r
    float dot = acc[0];
#pragma unroll
    for (int r = 1; r < kRowsPerIter; ++r) dot = lane == r ? acc[r] : dot;
    const int64_t row = r0 + lane;
    bool cand = false;
    float bn = 0.f;
    if (lane < kRowsPerIter && row < n) {
      bn = norm[row];
      float t = dot * __frcp_rn(bn);
      cand = !(t < thr);
    }
    unsigned m = __ballot_sync(0xffffffffu, cand);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      if (lane == src) {
        // the reference's operation order: dot / (|b| * |a|), IEEE fp32 (vo:182)
        float s = __fdiv_rn(dot, __fmul_rn(bn, an));
        uint64_t key = pack_key(s, (uint32_t)(row_base + row));
        if (key < below && key > mylist[k - 1]) topk_insert(mylist, k, key);
      }
      __syncwarp();
      thr = filter_threshold(mylist[k - 1], an);
    }
  }
  __syncthreads();
  // block merge: warp 0 selects the k best of kSingleWarps * k candidates
  if (wid == 0) {
    uint64_t prev = ~0ull;
    for (int r = 0; r < k; ++r) {
      uint64_t b = warp_next_best(lists, kSingleWarps * k, prev, lane);
      if (lane == 0) part[(size_t)blockIdx.x * k + r] = b;
      prev = b ? b : 0;  // once empty, stays empty
      if (b == 0) prev = 0;
    }
  }
  if constexpr (kFused) {
    __shared__ uint32_t s_last;
    if (threadIdx.x == 0) {
      __threadfence();                                  // this CTA's list is visible before the ticket is taken
      s_last = (atomicAdd(tail.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // every warp merges its share of the gridDim.x * k candidates, warp 0 merges the warps' lists
    uint64_t L[kLaneList];
    lane_list_clear(L);
    const int total = (int)gridDim.x * k;
    for (int i = threadIdx.x; i < total; i += kSingleThreads) lane_list_insert(L, k, __ldcg(&part[i]));
    uint64_t mine = warp_select_best(L, k, lane);
    if (lane < k) lists[wid * k + lane] = mine;
    __syncthreads();
    if (wid != 0) return;
    lane_list_clear(L);
    for (int i = lane; i < kSingleWarps * k; i += 32) lane_list_insert(L, k, lists[i]);
    mine = warp_select_best(L, k, lane);               // lane r < k: this rank's r-th best
    if (tail.world > 1) {
      const uint32_t par = tail.epoch & 1u;
      if (lane < k)
        for (int p = 0; p < tail.world; ++p)
          xchg_slot(tail.peer_bases[p], par, tail.world, tail.rank, tail.slot_stride)[lane] = mine;
      __threadfence_system();
      __syncwarp();
      if (lane < tail.world) {
        uint32_t* flags = reinterpret_cast<uint32_t*>(tail.peer_bases[lane]);
        st_release_sys(&flags[par * kXchgMaxWorld + tail.rank], tail.epoch);
        const uint32_t* own_flags = reinterpret_cast<const uint32_t*>(tail.peer_bases[tail.rank]);
        while (ld_acquire_sys(&own_flags[par * kXchgMaxWorld + lane]) != tail.epoch) __nanosleep(32);
      }
      __syncwarp();
      __threadfence_system();
      const uint64_t* gathered = xchg_slot(tail.peer_bases[tail.rank], par, tail.world, 0, tail.slot_stride);
      lane_list_clear(L);
      for (int i = lane; i < tail.world * k; i += 32) {
        const int p = i / k, j = i - p * k;
        lane_list_insert(L, k, __ldcg(&gathered[(size_t)p * tail.slot_stride + j]));
      }
      mine = warp_select_best(L, k, lane);
    }
    if (lane < k) {
      if (tail.out_idx) tail.out_idx[lane] = mine ? (int64_t)key_row(mine) : -1;
      if (tail.out_score) tail.out_score[lane] = mine ? key_score(mine) : 0.f;
      if (tail.out_key) tail.out_key[lane] = mine;
    }
    if (lane == 0) *tail.ticket = 0;
  }
}

// ---- two queries per bank pass -------------------------------------------------------------------------------------
// Between one query (the GEMV above) and the batches the tensor cores are built for lies the regime of a couple of
// concurrent questions: through the tcgen05 kernel two queries cost 3.7 ms over a 10M-row bank (its 256-query tile is
// mostly zero padding and is re-read for every bank tile), although the bank pass itself is the same 20.5 GB.  This
// kernel is the GEMV with NQ accumulators per row: the rows are streamed ONCE, each 128-bit load is expanded to fp32
// once and multiplied into NQ queries held in registers per 256-column chunk (the same FFMA order per query as the
// single-query kernel, hence the same bits).  Lists, thresholds and the paging cursor are per query.
// Measured on a 10M x 1024 bank (tools/batch_sweep.py): NQ = 2: 3.09 ms against 3.72 ms through tcgen05 (2.90 ms for
// one query); NQ = 4: 4.47 ms against 4.19 ms, NQ = 8: 9.4 ms against 4.22 ms -- the register budget (218 - 240) leaves
// one CTA of 8 warps per SM and the compiler sinks the prefetch loads into the arithmetic, so only NQ = 2 is
// dispatched here; three and more queries stay on the tensor-core kernel.
constexpr int kFewThreads = 256;
constexpr int kFewWarps = kFewThreads / 32;

template <int NQ>   // d == 1024 (four 256-column chunks); queries nq_real .. NQ-1 would be zero padding and produce nothing
__global__ void __launch_bounds__(kFewThreads, 1)
topk_few_kernel(const __nv_bfloat16* __restrict__ bank, const float* __restrict__ norm, int64_t n,
                const float* __restrict__ q /*[nq_real][1024]*/, int nq_real, int k, int64_t row_base,
                const uint64_t* __restrict__ after_key /*[nq_real] or null*/, uint64_t* __restrict__ part /*[grid][nq_real][k]*/) {
  constexpr int d = 1024, CH = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);                                   // [NQ][1024], permuted as above
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NQ * d * 4);     // [NQ][warps][k]
  __shared__ float s_an[NQ];
  __shared__ double s_part[kFewWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  for (int qi = 0; qi < NQ; ++qi) {
    double ss = 0.0;
    for (int e = threadIdx.x; e < d; e += kFewThreads) {
      const float v = qi < nq_real ? q[(size_t)qi * d + e] : 0.f;
      ss += (double)v * (double)v;
      const int c = e >> 8, l = (e >> 3) & 31, h = (e >> 2) & 1, j = e & 3;
      sq[qi * d + (((c * 2 + h) * 32) + l) * 4 + j] = v;
    }
    ss = warp_sum(ss);
    if (lane == 0) s_part[wid] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < kFewWarps; ++i) t += s_part[i];
      s_an[qi] = __fsqrt_rn((float)t);        // |a| as the reference forms it: fp32 sqrt of the sum of squares
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NQ * kFewWarps * k; i += kFewThreads) lists[i] = 0;
  __syncthreads();

  float an[NQ], thr[NQ];
  uint64_t below[NQ];
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) {
    an[qi] = s_an[qi];
    thr[qi] = -INFINITY;
    below[qi] = (after_key && qi < nq_real) ? after_key[qi] : ~0ull;
  }

  const int64_t ngroups = (n + kRowsPerIter - 1) / kRowsPerIter;
  const int64_t gwarp = (int64_t)blockIdx.x * kFewWarps + wid;
  const int64_t gstride = (int64_t)gridDim.x * kFewWarps;
  // the loads of group g + 1 are issued before the arithmetic of group g (register double buffer): with NQ
  // accumulators per row the kernel runs one CTA of 8 warps per SM, too few to hide the HBM latency by occupancy
  // alone (measured without the prefetch: 5.4 ms for 4 queries over 10M rows against 2.9 ms for one)
  auto load_group = [&](int64_t g, uint4 (&dst)[CH][kRowsPerIter]) {
    const int64_t r0 = g * kRowsPerIter;
#pragma unroll
    for (int r = 0; r < kRowsPerIter; ++r) {
      const int64_t row = r0 + r < n ? r0 + r : n - 1;
      const __nv_bfloat16* rp = bank + row * (int64_t)d + lane * 8;
#pragma unroll
      for (int c = 0; c < CH; ++c) dst[c][r] = ldg_stream(rp + c * 256);
    }
  };
  uint4 w[CH][kRowsPerIter];
  if (gwarp < ngroups) load_group(gwarp, w);
  for (int64_t g = gwarp; g < ngroups; g += gstride) {
    const int64_t r0 = g * kRowsPerIter;
    uint4 wn[CH][kRowsPerIter];
    load_group(g + gstride < ngroups ? g + gstride : g, wn);      // the last group is simply read again (L2 hit)
    float acc[NQ][kRowsPerIter];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r) acc[qi][r] = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float4 qa[NQ], qb[NQ];
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
        qa[qi] = *reinterpret_cast<const float4*>(&sq[qi * d + ((c * 2 + 0) * 32 + lane) * 4]);
        qb[qi] = *reinterpret_cast<const float4*>(&sq[qi * d + ((c * 2 + 1) * 32 + lane) * 4]);
      }
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r) {
        const uint4 x = w[c][r];
        const float x0 = bf16lo(x.x), x1 = bf16hi(x.x), x2 = bf16lo(x.y), x3 = bf16hi(x.y);
        const float x4 = bf16lo(x.z), x5 = bf16hi(x.z), x6 = bf16lo(x.w), x7 = bf16hi(x.w);
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          float a = acc[qi][r];
          a = fmaf(x0, qa[qi].x, a); a = fmaf(x1, qa[qi].y, a);
          a = fmaf(x2, qa[qi].z, a); a = fmaf(x3, qa[qi].w, a);
          a = fmaf(x4, qb[qi].x, a); a = fmaf(x5, qb[qi].y, a);
          a = fmaf(x6, qb[qi].z, a); a = fmaf(x7, qb[qi].w, a);
          acc[qi][r] = a;
        }
      }
    }
    const int64_t row = r0 + lane;                       // lane r owns row r0 + r
    const bool mine = lane < kRowsPerIter && row < n;
    const float bn = mine ? norm[row] : 1.f;
    const float rbn = __frcp_rn(bn);
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
      if (qi < nq_real) {                                // warp-uniform
#pragma unroll
        for (int r = 0; r < kRowsPerIter; ++r) acc[qi][r] = warp_sum(acc[qi][r]);
        float dot = acc[qi][0];
#pragma unroll
        for (int r = 1; r < kRowsPerIter; ++r) dot = lane == r ? acc[qi][r] : dot;
        const bool cand = mine && !(dot * rbn < thr[qi]);
        unsigned m = __ballot_sync(0xffffffffu, cand);
        if (m) {
          uint64_t* mylist = lists + ((size_t)qi * kFewWarps + wid) * k;
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            if (lane == src) {
              // the reference's operation order: dot / (|b| * |a|), IEEE fp32 (vo:182)
              const float sc = __fdiv_rn(dot, __fmul_rn(bn, an[qi]));
              const uint64_t key = pack_key(sc, (uint32_t)(row_base + row));
              if (key < below[qi] && key > mylist[k - 1]) topk_insert(mylist, k, key);
            }
            __syncwarp();
          }
          thr[qi] = filter_threshold(mylist[k - 1], an[qi]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r) w[c][r] = wn[c][r];
  }
  __syncthreads();
  // block merge: warp (qi mod warps) selects the k best of query qi's kFewWarps * k candidates
  for (int qi = wid; qi < nq_real; qi += kFewWarps) {
    uint64_t prev = ~0ull;
    for (int r = 0; r < k; ++r) {
      const uint64_t b = prev ? warp_next_best(lists + (size_t)qi * kFewWarps * k, kFewWarps * k, prev, lane) : 0;
      if (lane == 0) part[((size_t)blockIdx.x * nq_real + qi) * k + r] = b;
      prev = b;
    }
  }
}

static int few_grid_for(const void* fn, size_t smem) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kFewThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  return sm_count() * per_sm;
}

// Two queries of dimension 1024: one pass of the GEMV above + the per-query merge.  `part` needs
// topk_few_part_elems(nq, k) 8-byte words.  HIPPO_FEW_QUERIES=0 sends these batches through the tensor-core kernel.
bool topk_few_supported(int d, int nq) {
  static const bool off = getenv("HIPPO_FEW_QUERIES") && atoi(getenv("HIPPO_FEW_QUERIES")) == 0;
  return !off && d == 1024 && nq == 2;
}
size_t topk_few_part_elems(int nq, int k) { return (size_t)sm_count() * 2 * (size_t)nq * (size_t)k; }

// launches the GEMV only; *nparts_out receives the number of per-block lists in `part` ([nparts][nq][k])
hippo_status topk_few_launch(const void* bank, const float* norm, int64_t n, const float* q, int nq, int k,
                             int64_t row_base, const uint64_t* after_key, uint64_t* part, int* nparts_out,
                             cudaStream_t s) {
  constexpr int NQ = 2;
  HIPPO_REQUIRE(nq == NQ, "topk_few_launch: nq=%d", nq);
  const size_t smem = (size_t)NQ * 1024 * 4 + (size_t)NQ * kFewWarps * k * 8;
  const void* fn = (const void*)topk_few_kernel<NQ>;
  int grid = few_grid_for(fn, smem);
  const int64_t want = ((n + kRowsPerIter - 1) / kRowsPerIter + kFewWarps - 1) / kFewWarps;
  if (want < grid) grid = (int)(want < 1 ? 1 : want);
  topk_few_kernel<NQ><<<grid, kFewThreads, smem, s>>>((const __nv_bfloat16*)bank, norm, n, q, nq, k, row_base, after_key, part);
  HIPPO_CUDA(cudaGetLastError());
  *nparts_out = grid;
  return HIPPO_OK;
}

// keys [nparts, nq, k_in] -> best k per query; one warp per query.  ONE pass over the candidates: every lane
// keeps the best k of its share in a sorted register list, then the warp pops the k global maxima (the previous
// version rescanned all nparts * k_in candidates once per output rank: 76 us for 4,096 queries x 74 lists).
__global__ void __launch_bounds__(128) topk_merge_kernel(const uint64_t* __restrict__ keys, int nparts,
                                                         int nq, int k_in, int k,
                                                         int64_t* __restrict__ out_idx,
                                                         float* __restrict__ out_score,
                                                         uint64_t* __restrict__ out_key) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (qi >= nq) return;
  if (k <= kLaneList) {
    uint64_t L[kLaneList];
    lane_list_clear(L);
    const int total = nparts * k_in;
    for (int i = lane; i < total; i += 32) {
      const int p = i / k_in, j = i - p * k_in;
      const uint64_t c = keys[((size_t)p * nq + qi) * k_in + j];
      lane_list_insert(L, k, c);
    }
    const uint64_t mine = warp_select_best(L, k, lane);
    if (lane < k) {
      const size_t o = (size_t)qi * k + lane;
      if (out_idx) out_idx[o] = mine ? (int64_t)key_row(mine) : -1;
      if (out_score) out_score[o] = mine ? key_score(mine) : 0.f;
      if (out_key) out_key[o] = mine;
    }
    return;
  }
  // k beyond the register list (exact search merging many pages): selection by repeated scans
  uint64_t prev = ~0ull;
  for (int r = 0; r < k; ++r) {
    uint64_t best = 0;
    if (prev != 0) {
      for (int i = lane; i < nparts * k_in; i += 32) {
        int p = i / k_in, j = i - p * k_in;
        uint64_t c = keys[((size_t)p * nq + qi) * k_in + j];
        if (c < prev && c > best) best = c;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
    }
    if (lane == 0) {
      size_t o = (size_t)qi * k + r;
      if (out_idx) out_idx[o] = best ? (int64_t)key_row(best) : -1;
      if (out_score) out_score[o] = best ? key_score(best) : 0.f;
      if (out_key) out_key[o] = best;
    }
    prev = best;
  }
}

static size_t single_smem_bytes(int d, int k) {
  return (size_t)((d + 255) / 256 * 256) * 4 + (size_t)kSingleWarps * k * 8;
}
static int single_grid(int64_t n) {
  int64_t groups = (n + kRowsPerIter - 1) / kRowsPerIter;
  int64_t want = (groups + kSingleWarps - 1) / kSingleWarps;
  int64_t cap = (int64_t)sm_count() * 2;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace hippo

extern "C" {

size_t hippo_topk_single_workspace_bytes(int64_t n, int32_t d, int32_t k) {
  (void)d;
  int sms = hippo::sm_count();
  if (sms <= 0) sms = 148;
  (void)n;
  return hippo::align_up((size_t)sms * 2 * (size_t)(k > 0 ? k : 1) * 8 + 64, 256);   // per-block lists + the ticket
}

hippo_status hippo_topk_merge(const uint64_t* keys, int32_t nparts, int32_t nq, int32_t k_in,
                              int32_t k, int64_t* out_idx, float* out_score, uint64_t* out_key,
                              void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(nparts >= 1 && nq >= 0 && k_in >= 1 && k >= 1, "hippo_topk_merge: bad sizes");
  if (nq == 0) return HIPPO_OK;
  HIPPO_REQUIRE(keys != nullptr, "hippo_topk_merge: null keys");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  int blocks = (nq + 3) / 4;
  topk_merge_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(keys, nparts, nq, k_in, k, out_idx,
                                                              out_score, out_key);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

static hippo_status single_launch(const void* bank, const float* norm, int64_t n, int32_t d, const float* q, int32_t k,
                                  int64_t row_base, const uint64_t* after_key, void* ws, size_t ws_bytes,
                                  cudaStream_t s, const hippo::SingleTail* tail, int64_t* out_idx, float* out_score,
                                  uint64_t* out_key, const char* who) {
  using namespace hippo;
  HIPPO_REQUIRE(n >= 0 && d > 0 && d % 64 == 0, "%s: need d %% 64 == 0 (d=%d)", who, d);
  HIPPO_REQUIRE(k >= 1 && k <= HIPPO_TOPK_MAX, "%s: k=%d outside 1..%d", who, k, HIPPO_TOPK_MAX);
  HIPPO_REQUIRE(row_base >= 0 && row_base + n < 0xffffffffll, "%s: global row numbers must stay below 2^32-1", who);
  HIPPO_REQUIRE(q != nullptr && (n == 0 || (bank && norm)), "%s: null pointer", who);
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  const int grid = single_grid(n);
  if (ws == nullptr || ws_bytes < hippo_topk_single_workspace_bytes(n, d, k) || ((uintptr_t)ws & 255)) {
    set_error("%s: workspace of %zu bytes needed (256-byte aligned)", who, hippo_topk_single_workspace_bytes(n, d, k));
    return HIPPO_E_WORKSPACE;
  }
  uint64_t* part = (uint64_t*)ws;
  const size_t smem = single_smem_bytes(d, k);
  HIPPO_REQUIRE(smem <= 200 * 1024, "%s: d=%d too large", who, d);
  const __nv_bfloat16* b = (const __nv_bfloat16*)bank;
  if (tail == nullptr) {
    if (n == 0) {
      HIPPO_CUDA(cudaMemsetAsync(part, 0, (size_t)k * 8, s));
      return hippo_topk_merge(part, 1, 1, k, k, out_idx, out_score, out_key, (void*)s);
    }
    const SingleTail none{};
    if (d == 1024) {
      static bool attr4 = false;
      if (!attr4) { HIPPO_CUDA(cudaFuncSetAttribute(topk_single_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr4 = true; }
      topk_single_kernel<4, false><<<grid, kSingleThreads, smem, s>>>(b, norm, n, d, q, k, row_base, after_key, part, none);
    } else {
      static bool attr0 = false;
      if (!attr0) { HIPPO_CUDA(cudaFuncSetAttribute(topk_single_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr0 = true; }
      topk_single_kernel<0, false><<<grid, kSingleThreads, smem, s>>>(b, norm, n, d, q, k, row_base, after_key, part, none);
    }
    HIPPO_CUDA(cudaGetLastError());
    return hippo_topk_merge(part, grid, 1, k, k, out_idx, out_score, out_key, (void*)s);
  }
  // fused tail: the ticket sits behind the per-block lists (sized for 2 CTAs per SM); an empty shard still takes part
  // in the exchange (one CTA, every score list empty)
  SingleTail t = *tail;
  t.ticket = reinterpret_cast<uint32_t*>(part + (size_t)sm_count() * 2 * k);
  HIPPO_CUDA(cudaMemsetAsync(t.ticket, 0, 4, s));
  if (d == 1024) {
    static bool attr4f = false;
    if (!attr4f) { HIPPO_CUDA(cudaFuncSetAttribute(topk_single_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr4f = true; }
    topk_single_kernel<4, true><<<grid, kSingleThreads, smem, s>>>(b, norm, n, d, q, k, row_base, after_key, part, t);
  } else {
    static bool attr0f = false;
    if (!attr0f) { HIPPO_CUDA(cudaFuncSetAttribute(topk_single_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr0f = true; }
    topk_single_kernel<0, true><<<grid, kSingleThreads, smem, s>>>(b, norm, n, d, q, k, row_base, after_key, part, t);
  }
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

hippo_status hippo_topk_single(const void* bank, const float* norm, int64_t n, int32_t d,
                               const float* q, int32_t k, int64_t row_base,
                               const uint64_t* after_key, int64_t* out_idx, float* out_score,
                               uint64_t* out_key, void* ws, size_t ws_bytes, void* stream) {
  return single_launch(bank, norm, n, d, q, k, row_base, after_key, ws, ws_bytes, (cudaStream_t)stream, nullptr, out_idx,
                       out_score, out_key, "hippo_topk_single");
}

hippo_status hippo_topk_single_sharded(const void* bank, const float* norm, int64_t n, int32_t d, const float* q,
                                       int32_t k, int64_t row_base, const uint64_t* after_key,
                                       void* const* peer_bases, size_t buf_bytes, int32_t rank, int32_t world,
                                       uint32_t epoch, int64_t* out_idx, float* out_score, uint64_t* out_key,
                                       void* ws, size_t ws_bytes, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(world >= 1 && world <= kXchgMaxWorld && rank >= 0 && rank < world, "hippo_topk_single_sharded: bad rank / world");
  HIPPO_REQUIRE(epoch != 0, "hippo_topk_single_sharded: epoch 0 is the cleared state, start at 1");
  HIPPO_REQUIRE(world == 1 || peer_bases != nullptr, "hippo_topk_single_sharded: null peer table");
  HIPPO_REQUIRE(world == 1 || buf_bytes >= hippo_topk_exchange_bytes(world, 1, k),
                "hippo_topk_single_sharded: symmetric buffer of %zu bytes needed, got %zu",
                hippo_topk_exchange_bytes(world, 1, k), buf_bytes);
  SingleTail t{};
  t.peer_bases = (unsigned char* const*)peer_bases;
  t.slot_stride = world > 1 ? (buf_bytes - kXchgHeader) / ((size_t)2 * world * sizeof(uint64_t)) : 0;
  t.rank = rank;
  t.world = world;
  t.epoch = epoch;
  t.out_idx = out_idx;
  t.out_score = out_score;
  t.out_key = out_key;
  return single_launch(bank, norm, n, d, q, k, row_base, after_key, ws, ws_bytes, (cudaStream_t)stream, &t, out_idx,
                       out_score, out_key, "hippo_topk_single_sharded");
}

}  // extern "C"
